// cge_ingest.cpp -- SURVEY.md section 8(f) row F3: the whitespace-delimited numeric tables the
// scorer is fed from (edgelist, communities, embedding; /root/reference/src/auxilary.jl:86-168 reads
// them with DelimitedFiles.readdlm(fn, Float64 | Int)).  At 1M vertices x 128 dimensions the
// embedding is ~1-2.5 GB of text and its parse dominates the host time of a run; this reader
// maps the file, splits it at line boundaries and parses the pieces on all host cores straight
// into the caller's matrix (any strides: Julia's column-major Matrix{Float64} or a C array).
//
// Semantics kept from the typed readdlm call the reference makes:
//   * cells are separated by runs of blanks / tabs, rows by '\n' ('\r' before it is ignored);
//   * blank lines are skipped (skipblanks = true);
//   * every row must have the column count of the first row, every cell must be a number --
//     otherwise the call fails (the reference relies on that failure to detect the node2vec
//     header line, auxilary.jl:150-155);
//   * numbers are converted with correct rounding (std::from_chars / strtod), like Julia's parser.
// Host-only code: no CUDA call, usable without a GPU.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/cge_b200.h"

namespace cge {
void set_last_error(const std::string &msg);  // cge_host.cu
}

namespace {

struct Mapped {
    const char *p = nullptr;
    size_t len = 0;
    int fd = -1;
    ~Mapped() {
        if (p && len) munmap(const_cast<char *>(p), len);
        if (fd >= 0) close(fd);
    }
};

int fail(int code, const std::string &msg) {
    cge::set_last_error(msg);
    return code;
}

int map_file(const char *path, Mapped &m) {
    m.fd = open(path, O_RDONLY);
    if (m.fd < 0) return fail(CGE_B200_ERR_ARG, std::string(path) + " is not a file");
    struct stat st;
    if (fstat(m.fd, &st) != 0 || !S_ISREG(st.st_mode))
        return fail(CGE_B200_ERR_ARG, std::string(path) + " is not a file");
    m.len = (size_t)st.st_size;
    if (m.len == 0) return 0;
    void *p = mmap(nullptr, m.len, PROT_READ, MAP_PRIVATE, m.fd, 0);
    if (p == MAP_FAILED) {
        m.len = 0;
        return fail(CGE_B200_ERR_OOM, std::string("mmap(") + path + ") failed");
    }
    madvise(p, m.len, MADV_SEQUENTIAL);
    m.p = (const char *)p;
    return 0;
}

inline bool is_blank(char c) { return c == ' ' || c == '\t' || c == '\r'; }

// [b, e) = one line without its '\n'; true when it holds no cell
inline bool blank_line(const char *b, const char *e) {
    for (; b < e; ++b)
        if (!is_blank(*b)) return false;
    return true;
}

// offset of the first byte after `skip` lines
size_t skip_lines(const Mapped &m, int64_t skip) {
    size_t pos = 0;
    while (skip > 0 && pos < m.len) {
        const void *nl = memchr(m.p + pos, '\n', m.len - pos);
        pos = nl ? (size_t)((const char *)nl - m.p) + 1 : m.len;
        --skip;
    }
    return pos;
}

// number of cells of a line
int64_t count_cells(const char *b, const char *e) {
    int64_t n = 0;
    while (b < e) {
        while (b < e && is_blank(*b)) ++b;
        if (b == e) break;
        ++n;
        while (b < e && !is_blank(*b)) ++b;
    }
    return n;
}

// one cell -> double; false when the token is not a number in full
bool parse_cell(const char *b, const char *e, double &v) {
    if (b < e && *b == '+') ++b;  // from_chars rejects an explicit plus sign
    auto r = std::from_chars(b, e, v);
    if (r.ec == std::errc() && r.ptr == e) return true;
    // hexadecimal floats, overflow to Inf and the like: the C library decides
    char tmp[64];
    const size_t n = (size_t)(e - b);
    if (n == 0 || n >= sizeof(tmp)) return false;
    memcpy(tmp, b, n);
    tmp[n] = 0;
    char *end = nullptr;
    v = strtod(tmp, &end);
    return end == tmp + n;
}

struct Piece {
    size_t begin, end;   // byte range, begins at a line start, ends after a '\n' (or at EOF)
    int64_t rows = 0;    // non-blank lines inside
    int64_t row0 = 0;    // index of its first row in the table
};

// cut [start, len) into ~n pieces at line boundaries
std::vector<Piece> cut(const Mapped &m, size_t start, int n) {
    std::vector<Piece> out;
    size_t pos = start;
    const size_t total = m.len - start;
    for (int i = 0; i < n && pos < m.len; ++i) {
        size_t want = start + (size_t)((double)total * (i + 1) / n);
        if (i == n - 1 || want >= m.len) want = m.len;
        if (want < pos) want = pos;
        if (want < m.len) {
            const void *nl = memchr(m.p + want, '\n', m.len - want);
            want = nl ? (size_t)((const char *)nl - m.p) + 1 : m.len;
        }
        if (want > pos) out.push_back({pos, want});
        pos = want;
    }
    return out;
}

template <typename F>
void for_lines(const Mapped &m, const Piece &pc, F &&f) {
    size_t pos = pc.begin;
    while (pos < pc.end) {
        const void *nl = memchr(m.p + pos, '\n', pc.end - pos);
        const size_t e = nl ? (size_t)((const char *)nl - m.p) : pc.end;
        if (!f(m.p + pos, m.p + e)) return;
        pos = e + 1;
    }
}

int n_workers(int32_t n_threads, size_t bytes) {
    int n = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    if (n < 1) n = 1;
    const size_t by_size = bytes / (1 << 16) + 1;  // no point in a thread per few KB
    return (int)std::min<size_t>((size_t)n, by_size);
}

template <typename F>
void run_parallel(int n, F &&f) {
    if (n <= 0) return;
    std::vector<std::thread> th;
    for (int i = 1; i < n; ++i) th.emplace_back([&f, i]() { f(i); });
    f(0);
    for (auto &t : th) t.join();
}

}  // namespace

extern "C" {

int cge_b200_table_dims(const char *path, int64_t skip_rows, int32_t n_threads, int64_t *rows,
                        int64_t *cols) {
    if (!path || skip_rows < 0 || !rows || !cols) return fail(CGE_B200_ERR_ARG, "NULL argument");
    Mapped m;
    if (int rc = map_file(path, m)) return rc;
    const size_t start = skip_lines(m, skip_rows);
    *rows = 0;
    *cols = 0;
    if (start >= m.len) return 0;
    auto pieces = cut(m, start, n_workers(n_threads, m.len - start));
    run_parallel((int)pieces.size(), [&](int i) {
        int64_t r = 0;
        for_lines(m, pieces[i], [&](const char *b, const char *e) {
            if (!blank_line(b, e)) ++r;
            return true;
        });
        pieces[i].rows = r;
    });
    for (auto &pc : pieces) *rows += pc.rows;
    // the column count is that of the first row
    for (auto &pc : pieces) {
        if (pc.rows == 0) continue;
        for_lines(m, pc, [&](const char *b, const char *e) {
            if (blank_line(b, e)) return true;
            *cols = count_cells(b, e);
            return false;
        });
        break;
    }
    return 0;
}

int cge_b200_read_table(const char *path, int64_t skip_rows, int32_t n_threads, int64_t rows,
                        int64_t cols, int64_t row_stride, int64_t col_stride, double *out) {
    if (!path || skip_rows < 0 || rows < 0 || cols < 0 || (rows > 0 && cols > 0 && !out))
        return fail(CGE_B200_ERR_ARG, "bad read_table argument");
    Mapped m;
    if (int rc = map_file(path, m)) return rc;
    const size_t start = skip_lines(m, skip_rows);
    std::vector<Piece> pieces;
    if (start < m.len) pieces = cut(m, start, n_workers(n_threads, m.len - start));
    // pass 1: rows per piece -> first row index of every piece
    run_parallel((int)pieces.size(), [&](int i) {
        int64_t r = 0;
        for_lines(m, pieces[i], [&](const char *b, const char *e) {
            if (!blank_line(b, e)) ++r;
            return true;
        });
        pieces[i].rows = r;
    });
    int64_t total = 0;
    for (auto &pc : pieces) {
        pc.row0 = total;
        total += pc.rows;
    }
    if (total != rows)
        return fail(CGE_B200_ERR_ARG, std::string(path) + ": " + std::to_string(total) +
                                          " rows, the caller expects " + std::to_string(rows));
    // pass 2: parse
    std::atomic<int64_t> bad_row{-1};
    std::atomic<int> bad_kind{0};  // 1 = column count, 2 = not a number
    run_parallel((int)pieces.size(), [&](int i) {
        int64_t r = pieces[i].row0;
        for_lines(m, pieces[i], [&](const char *b, const char *e) {
            if (blank_line(b, e)) return true;
            if (bad_row.load(std::memory_order_relaxed) >= 0) return false;
            int64_t c = 0;
            int kind = 0;
            while (b < e) {
                while (b < e && is_blank(*b)) ++b;
                if (b == e) break;
                const char *t = b;
                while (b < e && !is_blank(*b)) ++b;
                if (c >= cols) {
                    kind = 1;
                    break;
                }
                double v;
                if (!parse_cell(t, b, v)) {
                    kind = 2;
                    break;
                }
                out[r * row_stride + c * col_stride] = v;
                ++c;
            }
            if (!kind && c != cols) kind = 1;
            if (kind) {
                int64_t expect = -1;
                if (bad_row.compare_exchange_strong(expect, r)) bad_kind.store(kind);
                return false;
            }
            ++r;
            return true;
        });
    });
    if (bad_row.load() >= 0)
        return fail(CGE_B200_ERR_ARG,
                    std::string(path) + ": row " + std::to_string(bad_row.load() + 1 + skip_rows) +
                        (bad_kind.load() == 1 ? " does not have " + std::to_string(cols) + " columns"
                                              : " holds a cell that is not a number"));
    return 0;
}

}  // extern "C"
