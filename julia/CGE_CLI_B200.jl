# Runs the reference's own CLI script on the B200 scorer without copying it: the text of
# example/CGE_CLI.jl is read from the CGE.jl checkout, its `using CGE` line is dropped and the
# rest is evaluated after `using .CGEB200`, whose wGCL / wGCL_directed / parseargs / landmarks are
# then the names the script resolves.  Every flag of the reference CLI keeps working.
#
#   julia CGE_CLI_B200.jl -g G.edgelist -c G.ecg -e G.embedding -l 200 --seed 42
#
# The script is located through ENV["CGE_REFERENCE_CLI"] or pathof(CGE).
include(joinpath(@__DIR__, "CGEB200.jl"))
using .CGEB200
import CGE

cli = get(ENV, "CGE_REFERENCE_CLI",
          normpath(joinpath(dirname(pathof(CGE)), "..", "example", "CGE_CLI.jl")))
isfile(cli) || error("reference CLI script not found: $cli (set CGE_REFERENCE_CLI)")
body = replace(read(cli, String), r"^\s*using\s+CGE\s*$"m => "")
include_string(Main, body, cli)
