// csv_codec.cu -- lib/csv.h: the reference's checkpoint / data format, byte-compatible and parallel.
//
// The reference writes every parameter tensor as text, `fprintf(f, "%f,", v)` per value with a newline after each row
// (lib/csv.c:56-67), and reads it back one fgetc at a time into atof (lib/csv.c:28-54): 1.9 MB for the MLP's first layer,
// ~450 MB for the U-Net.  Here the file is read (or produced) as one buffer, split at value boundaries across the host
// cores, and converted with exact integer arithmetic:
//   format : a float is m * 2^e with a 24-bit m, so round_half_even(v * 10^6) is a 64-bit shift with a remainder test --
//            the same digits glibc's correctly rounded printf("%f") prints, without its big-number path;
//   parse  : std::from_chars<double> (correctly rounded, as strtod/atof) then the reference's implicit double -> float
//            conversion; tokens that are not plain decimals (spaces, inf, nan, hex) go through strtod itself.
// Same tokenisation as the reference, quirks included: a value ends at ',' (an empty field is 0), at '\n' only if the field
// is not empty, '\r' is skipped, characters after the last terminator are dropped, *num_values is the number of commas.
// Device-resident tensors are staged through pinned memory (bla_csv_save_device / bla_csv_load_device).
#include <algorithm>
#include <atomic>
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include "../../include/bla.h"
#include "../../include/lib/csv.h"
#include "runtime.h"

namespace {

int worker_count(size_t bytes) {
    if (bytes < (1u << 20)) return 1;
    unsigned hc = std::thread::hardware_concurrency();
    int n = hc ? (int)hc : 4;
    if (n > 32) n = 32;
    const int by_size = (int)(bytes >> 19);   // at least 512 KB of text per thread
    return std::max(1, std::min(n, by_size));
}

template <class Fn>
void parallel_for(int n, Fn fn) {
    if (n <= 1) { fn(0); return; }
    std::vector<std::thread> th;
    th.reserve(n - 1);
    for (int i = 1; i < n; ++i) th.emplace_back(fn, i);
    fn(0);
    for (auto& t : th) t.join();
}

// ---- "%f," ----------------------------------------------------------------------------------------------------------
// "00".."99"
struct DigitPairs {
    char d[200];
    constexpr DigitPairs() : d() {
        for (int i = 0; i < 100; ++i) { d[2 * i] = (char)('0' + i / 10); d[2 * i + 1] = (char)('0' + i % 10); }
    }
};
constexpr DigitPairs kPairs{};

inline char* put_u64(char* p, uint64_t v) {
    if (v < 10) { *p++ = (char)('0' + v); return p; }   // weights: almost always "0"
    char tmp[20];
    int n = 20;
    while (v >= 100) { n -= 2; memcpy(tmp + n, kPairs.d + 2 * (v % 100), 2); v /= 100; }
    if (v >= 10) { n -= 2; memcpy(tmp + n, kPairs.d + 2 * v, 2); } else tmp[--n] = (char)('0' + v);
    memcpy(p, tmp + n, 20 - n);
    return p + (20 - n);
}

// appends printf("%f,", (double)x) to p (at most 48 bytes for |x| < 2^63; larger magnitudes, inf and nan go through snprintf)
inline char* format_value(char* p, float x) {
    uint32_t bits;
    memcpy(&bits, &x, 4);
    const uint32_t E = (bits >> 23) & 0xFF;
    uint64_t M = bits & 0x7FFFFFu;
    if (E == 0xFF || E >= 150 + 39) {   // inf / nan / >= 2^62: glibc's general path
        return p + snprintf(p, 64, "%f,", (double)x);
    }
    int e;
    if (E == 0) e = -149; else { M |= 0x800000u; e = (int)E - 150; }
    if (bits >> 31) *p++ = '-';
    uint64_t N;   // round_half_even(|x| * 10^6)
    if (e >= 0) {
        N = 0;     // integer: handled below without the 10^6 factor
        p = put_u64(p, M << e);
        memcpy(p, ".000000,", 8);
        return p + 8;
    }
    const int s = -e;
    const uint64_t P = M * 1000000ull;   // < 2^44
    if (s >= 45) {
        N = 0;                            // P < 2^44 <= half of 2^s
    } else {
        N = P >> s;
        const uint64_t rem = P & ((1ull << s) - 1), half = 1ull << (s - 1);
        if (rem > half || (rem == half && (N & 1))) ++N;
    }
    p = put_u64(p, N / 1000000ull);
    uint32_t frac = (uint32_t)(N % 1000000ull);
    *p++ = '.';
    memcpy(p + 4, kPairs.d + 2 * (frac % 100), 2); frac /= 100;
    memcpy(p + 2, kPairs.d + 2 * (frac % 100), 2); frac /= 100;
    memcpy(p, kPairs.d + 2 * frac, 2);
    p[6] = ',';
    return p + 7;
}

// text of rows [row0, row1); the buffer is sized for the worst case but never initialised, so only the pages the text
// reaches are ever touched
struct TextPart {
    std::unique_ptr<char[]> buf;
    size_t size = 0;
};

void format_rows(const float* data, int cols, size_t row0, size_t row1, TextPart& out) {
    out.buf.reset(new char[(row1 - row0) * ((size_t)cols * 50 + 1) + 64]);
    char* p = out.buf.get();
    for (size_t r = row0; r < row1; ++r) {
        const float* v = data + r * cols;
        for (int c = 0; c < cols; ++c) p = format_value(p, v[c]);
        *p++ = '\n';
    }
    out.size = (size_t)(p - out.buf.get());
}

int format_workers(size_t rows, int cols) {
    return std::max(1, std::min(worker_count(rows * (size_t)cols * 10), (int)std::max<size_t>(rows, 1)));
}

// ---- atof ------------------------------------------------------------------------------------------------------------
// What "%f" writes -- [-]digits[.digits], at most 15 digits in all -- read in one walk.  The digits are an integer below 2^53
// and the power of ten is exact, so one IEEE division gives the correctly rounded double strtod would return.  Returns the
// first character not consumed; *ok is false when the text walked is not of that form (the caller then takes the general path).
inline const char* walk_fixed(const char* q, const char* e, double* v, bool* ok) {
    static const double kPow10[16] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15};
    const bool neg = q < e && *q == '-';
    q += neg;
    uint64_t m = 0;
    const char* first = q;
    for (; q < e && (unsigned)(*q - '0') < 10u; ++q) m = m * 10 + (unsigned)(*q - '0');
    int digits = (int)(q - first), frac = 0;
    if (q < e && *q == '.') {
        const char* dot = ++q;
        for (; q < e && (unsigned)(*q - '0') < 10u; ++q) m = m * 10 + (unsigned)(*q - '0');
        frac = (int)(q - dot);
        digits += frac;
    }
    *ok = digits > 0 && digits <= 15;
    const double x = (double)m / kPow10[frac & 15];
    *v = neg ? -x : x;
    return q;
}

inline float parse_token(const char* b, const char* e) {
    // the reference: atof on the field (leading blanks skipped, trailing junk ignored, "" -> 0)
    {
        double v;
        bool ok;
        if (walk_fixed(b, e, &v, &ok) == e && ok) return (float)v;
    }
    const char* q = b;
    bool plain = true;
    if (q < e && *q == '+') { ++q; plain = !(q < e && (*q == '+' || *q == '-')); }   // strtod("+-1") is 0, not -1
    double v = 0.0;
    if (plain) {
        auto r = std::from_chars(q, e, v);
        if (r.ec == std::errc() && r.ptr == e) return (float)v;
    }
    char tmp[1100];
    size_t n = std::min((size_t)(e - b), sizeof(tmp) - 1);
    memcpy(tmp, b, n);
    tmp[n] = '\0';
    return (float)atof(tmp);
}

struct Scan { size_t values = 0, commas = 0; };

// Walks [b, e) with the reference's state machine, field by field, storing the values in out (room for e - b of them: a
// value needs at least its terminator).  Chunk boundaries are placed right after a terminator, so b starts a field.  A field
// written by "%f" is read and delimited in the same walk; anything else (empty, exponent, blanks, '\r', text) is delimited
// first and goes through parse_token.
Scan scan(const char* b, const char* e, float* out) {
    Scan s;
    const char* tok = b;
    while (tok < e) {
        double v;
        bool ok;
        const char* q = walk_fixed(tok, e, &v, &ok);
        if (ok && q < e && (*q == ',' || *q == '\n')) {
            out[s.values++] = (float)v;
            s.commas += *q == ',';
            tok = q + 1;
            continue;
        }
        const char* p = q;
        while (p < e && *p != ',' && *p != '\n') ++p;
        if (p == e) break;                        // no terminator: the reference drops what follows the last one
        const char c = *p;
        if (memchr(tok, '\r', (size_t)(p - tok))) {   // rare: strip the '\r's into a scratch copy
            char tmp[1100];
            size_t n = 0;
            for (const char* r = tok; r < p && n < sizeof(tmp) - 1; ++r) if (*r != '\r') tmp[n++] = *r;
            if (c == ',' || n > 0) out[s.values++] = parse_token(tmp, tmp + n);
        } else if (c == ',' || p > tok) {
            out[s.values++] = parse_token(tok, p);
        }
        s.commas += c == ',';
        tok = p + 1;
    }
    return s;
}

float* parse_buffer(const char* text, size_t len, int* num_values, size_t* count_out) {
    const int T = worker_count(len);
    std::vector<size_t> cut(T + 1, 0);
    cut[T] = len;
    for (int i = 1; i < T; ++i) {   // boundaries right after a terminator
        size_t p = len / T * i;
        while (p < len && text[p] != ',' && text[p] != '\n') ++p;
        cut[i] = std::min(len, p + 1);
    }
    for (int i = 1; i <= T; ++i) cut[i] = std::max(cut[i], cut[i - 1]);
    // one pass: every worker parses its chunk into a buffer of its own sized for the worst case (never initialised, so only the
    // pages the values reach are touched); the values are then gathered into one array
    std::vector<Scan> counts(T);
    std::vector<float*> local(T, nullptr);
    std::vector<char> oom(T, 0);
    parallel_for(T, [&](int i) {
        const size_t bytes = cut[i + 1] - cut[i];
        local[i] = (float*)malloc((bytes + 1) * sizeof(float));
        if (!local[i]) { oom[i] = 1; return; }
        counts[i] = scan(text + cut[i], text + cut[i + 1], local[i]);
    });
    size_t total = 0, commas = 0;
    std::vector<size_t> first(T);
    for (int i = 0; i < T; ++i) { first[i] = total; total += counts[i].values; commas += counts[i].commas; }
    // the reference sizes its buffer by the comma count (lib/csv.c:29-35) and overflows it when rows lack the trailing comma
    // (SURVEY D8); here the buffer always holds every value
    float* out = T == 1 && local[0] ? local[0] : (float*)malloc(std::max<size_t>(std::max(total, commas), 1) * sizeof(float));
    if (!out || std::find(oom.begin(), oom.end(), 1) != oom.end()) bla::die("bla: out of memory parsing a CSV of %zu bytes, exiting", len);
    if (out != local[0]) {
        parallel_for(T, [&](int i) { memcpy(out + first[i], local[i], counts[i].values * sizeof(float)); free(local[i]); });
    }
    if (num_values) *num_values = (int)commas;
    if (count_out) *count_out = total;
    return out;
}

// The whole file as one buffer (malloc'd, uninitialised until read).  A regular file is read by all workers at once, each its
// slice with pread (page-cache copies scale with the cores); anything else (a pipe) is read to its end through the FILE.
struct Text {
    char* data = nullptr;
    size_t size = 0;
    ~Text() { free(data); }
};

void slurp(FILE* f, Text& t) {
    rewind(f);
    struct stat st;
    const int fd = fileno(f);
    if (fd >= 0 && fstat(fd, &st) == 0 && S_ISREG(st.st_mode) && st.st_size > 0) {
        const size_t n = (size_t)st.st_size;
        t.data = (char*)malloc(n);
        if (!t.data) bla::die("bla: out of memory reading a CSV of %zu bytes, exiting", n);
        const int T = worker_count(n);
        std::vector<size_t> got(T, 0);
        parallel_for(T, [&](int i) {
            size_t at = n / T * i;
            const size_t end = i + 1 == T ? n : n / T * (i + 1);
            while (at < end) {
                const ssize_t r = pread(fd, t.data + at, end - at, (off_t)at);
                if (r <= 0) break;
                at += (size_t)r;
                got[i] += (size_t)r;
            }
        });
        // a file cut short under the reader: keep the leading part that did arrive
        for (int i = 0; i < T; ++i) {
            t.size += got[i];
            if (got[i] != (i + 1 == T ? n : n / T * (i + 1)) - n / T * i) break;
        }
        return;
    }
    size_t cap = 1 << 16, got;
    t.data = (char*)malloc(cap);
    while (t.data && (got = fread(t.data + t.size, 1, cap - t.size, f)) > 0) {
        t.size += got;
        if (t.size == cap) t.data = (char*)realloc(t.data, cap *= 2);
    }
    if (!t.data) bla::die("bla: out of memory reading a CSV, exiting");
}

// The rows are cut into more parts than workers; the workers format parts in order while this thread writes every part as soon
// as it is ready and frees it, so formatting hides behind the file system (which takes ~3x longer) and at most a few parts
// are in memory at once.  Plain write(): also right for a pipe or /dev/stdout.
void write_text(const char* filepath, const float* data, int cols, size_t rows) {
    const int fd = open(filepath, O_WRONLY | O_CREAT | O_TRUNC, 0666);
    if (fd < 0) bla::die("bla: cannot open CSV file %s for writing, exiting", filepath);
    const int T = format_workers(rows, cols);
    const int P = T == 1 ? 1 : (int)std::min<size_t>(rows, (size_t)T * 4);
    std::vector<TextPart> parts(P);
    std::vector<char> ready(P, 0);
    std::mutex mu;
    std::condition_variable cv;
    std::atomic<int> next{0};
    auto work = [&] {
        for (int k; (k = next.fetch_add(1)) < P;) {
            format_rows(data, cols, rows * k / P, rows * (k + 1) / P, parts[k]);
            { std::lock_guard<std::mutex> g(mu); ready[k] = 1; }
            cv.notify_one();
        }
    };
    std::vector<std::thread> th;
    if (T == 1) work();
    else for (int i = 0; i < T; ++i) th.emplace_back(work);
    bool ok = true;
    for (int k = 0; k < P; ++k) {
        { std::unique_lock<std::mutex> g(mu); cv.wait(g, [&] { return ready[k] != 0; }); }
        const char* q = parts[k].buf.get();
        size_t left = ok ? parts[k].size : 0;
        while (left) {
            const ssize_t w = write(fd, q, left);
            if (w <= 0) { ok = false; break; }
            q += w; left -= (size_t)w;
        }
        parts[k].buf.reset();
    }
    for (auto& t : th) t.join();
    if (close(fd) != 0 || !ok) bla::die("bla: short write to %s, exiting", filepath);
}

}  // namespace

extern "C" {

// lib/csv.c:18-25
float* read_csv_contents(const char* filepath) {
    FILE* f = fopen(filepath, "r");
    if (!f) bla::die("bla: cannot open CSV file %s, exiting", filepath);
    return read_csv_contents_file(f, nullptr);
}

// lib/csv.c:28-54.  Side-effect (as the reference): closes f.  The result is malloc'd: callers free() it or hand it to make_matrix.
float* read_csv_contents_file(FILE* f, int* num_values) {
    Text text;
    slurp(f, text);
    fclose(f);
    return parse_buffer(text.data, text.size, num_values, nullptr);
}

// lib/csv.c:56-67
void write_csv_contents(const char* filepath, float* data, int cols, int rows) {
    if (cols <= 0 || rows <= 0) { FILE* f = fopen(filepath, "w"); if (f) fclose(f); return; }
    write_text(filepath, data, cols, (size_t)rows);
}

// lib/csv.c:70-89
int count_num_lines(FILE* f) {
    char buffer[65536];
    int count = 0;
    for (;;) {
        size_t got = fread(buffer, 1, sizeof(buffer), f);
        if (ferror(f)) return -1;
        for (size_t i = 0; i < got; ++i) count += buffer[i] == '\n';
        if (feof(f)) break;
    }
    return count;
}

// ---- additive (include/bla.h "CSV checkpoint codec") --------------------------------------------------------------
size_t bla_csv_parse(const char* text, size_t len, float** values_out) {
    size_t n = 0;
    *values_out = parse_buffer(text, len, nullptr, &n);
    return n;
}

size_t bla_csv_format(const float* data, int cols, size_t rows, char* out, size_t cap) {
    const int T = format_workers(rows, cols);
    std::vector<TextPart> parts(T);
    parallel_for(T, [&](int i) { format_rows(data, cols, rows * i / T, rows * (i + 1) / T, parts[i]); });
    std::vector<size_t> at(T + 1, 0);
    for (int i = 0; i < T; ++i) at[i + 1] = at[i] + parts[i].size;
    if (out && at[T] <= cap) parallel_for(T, [&](int i) { memcpy(out + at[i], parts[i].buf.get(), parts[i].size); });
    return at[T];
}

// a [rows x cols] tensor in any memory (device, managed, host) -> the reference's CSV
void bla_csv_save(const char* filepath, const float* data, int cols, size_t rows) {
    using namespace bla;
    const size_t n = rows * (size_t)cols;
    if (classify(data) == kDevice) {
        float* pin = (float*)pool_alloc(kPinned, n * sizeof(float));
        BLA_CUDA(cudaMemcpyAsync(pin, data, n * sizeof(float), cudaMemcpyDeviceToHost, rt().stream));
        BLA_CUDA(cudaStreamSynchronize(rt().stream));
        rt().d2h_bytes += n * sizeof(float);
        write_text(filepath, pin, cols, rows);
        pool_free(pin);
    } else {
        if (rt_initialised()) BLA_CUDA(cudaStreamSynchronize(rt().stream));
        write_text(filepath, data, cols, rows);
    }
}

// the reference's CSV -> `count` floats at dst (device, managed or host); exits if the file holds fewer values
void bla_csv_load(const char* filepath, float* dst, size_t count) {
    using namespace bla;
    FILE* f = fopen(filepath, "r");
    if (!f) die("bla: cannot open CSV file %s, exiting", filepath);
    Text text;
    slurp(f, text);
    fclose(f);
    size_t n = 0;
    float* v = parse_buffer(text.data, text.size, nullptr, &n);
    if (n < count) die("bla: CSV file %s holds %zu values, %zu expected, exiting", filepath, n, count);
    if (classify(dst) == kDevice) {
        float* pin = (float*)pool_alloc(kPinned, count * sizeof(float));
        memcpy(pin, v, count * sizeof(float));
        BLA_CUDA(cudaMemcpyAsync(dst, pin, count * sizeof(float), cudaMemcpyHostToDevice, rt().stream));
        BLA_CUDA(cudaStreamSynchronize(rt().stream));
        rt().h2d_bytes += count * sizeof(float);
        pool_free(pin);
    } else {
        memcpy(dst, v, count * sizeof(float));
    }
    free(v);
}

}  // extern "C"
