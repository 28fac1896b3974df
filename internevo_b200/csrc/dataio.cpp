// Host-side data IO helpers for the tokenised-corpus format ({"tokens": [..]} JSON lines in a .bin file + a .bin.meta table of
// (byte offset, token count) pairs; reference internlm/data/tokenized/single_dataset.py:18-117, tools/tokenizer.py).
//
//   b200_scan_jsonl    one pass over a .bin file -> (offset, n_tokens) of every line: builds the .meta table of a corpus without
//                      decoding a single JSON document in Python (the table of a 100 GB shard is minutes of json.loads otherwise)
//   b200_parse_tokens  the token list of ONE line straight into an int64 buffer (the per-sample work of every data-loader worker)
//
// Plain C ABI (loaded with ctypes, no Python / torch headers): built by csrc/build.py into internevo_b200/_dataio.so.  Both
// functions accept exactly the shape the tokenisers write - one object whose only member is "tokens": [ints] - and report anything
// else as -1 so the caller falls back to the JSON decoder for that line.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace {

inline const char* skip_ws(const char* p, const char* e) {
    while (p < e && (*p == ' ' || *p == '\t' || *p == '\r' || *p == '\n')) ++p;
    return p;
}

// Parses {"tokens": [a, b, ...]} in [p, e).  out == nullptr: count only.  Returns the token count, -1 if the text has another shape,
// -2 if out is too small.
int64_t parse_line(const char* p, const char* e, int64_t* out, int64_t cap) {
    p = skip_ws(p, e);
    if (p >= e || *p != '{') return -1;
    p = skip_ws(p + 1, e);
    static const char key[] = "\"tokens\"";
    const int64_t klen = sizeof(key) - 1;
    if (e - p < klen || std::memcmp(p, key, klen) != 0) return -1;
    p = skip_ws(p + klen, e);
    if (p >= e || *p != ':') return -1;
    p = skip_ws(p + 1, e);
    if (p >= e || *p != '[') return -1;
    p = skip_ws(p + 1, e);
    int64_t n = 0;
    if (p < e && *p == ']') {
        ++p;
    } else {
        for (;;) {
            bool neg = false;
            if (p < e && *p == '-') { neg = true; ++p; }
            if (p >= e || *p < '0' || *p > '9') return -1;
            int64_t v = 0;
            int digits = 0;
            while (p < e && *p >= '0' && *p <= '9') {
                v = v * 10 + (*p - '0');
                ++p;
                if (++digits > 18) return -1;
            }
            if (p < e && (*p == '.' || *p == 'e' || *p == 'E')) return -1;   // not an integer: leave it to the JSON decoder
            if (out != nullptr) {
                if (n >= cap) return -2;
                out[n] = neg ? -v : v;
            }
            ++n;
            p = skip_ws(p, e);
            if (p >= e) return -1;
            if (*p == ',') { p = skip_ws(p + 1, e); continue; }
            if (*p == ']') { ++p; break; }
            return -1;
        }
    }
    p = skip_ws(p, e);
    if (p >= e || *p != '}') return -1;
    p = skip_ws(p + 1, e);
    return p == e ? n : -1;
}

}  // namespace

extern "C" {

// Token count of the JSON line in data[0, len) and, if out != NULL, its tokens.  -1: not the plain {"tokens": [...]} form.
int64_t b200_parse_tokens(const char* data, int64_t len, int64_t* out, int64_t cap) {
    return parse_line(data, data + len, out, cap);
}

// (offset, n_tokens) of every line of the file into out[2 * i], out[2 * i + 1] (up to cap lines; out == NULL: count lines only).
// Returns the number of lines, -1 if the file cannot be read, -(line + 2) for the first line (0-based) that is not in the plain form.
int64_t b200_scan_jsonl(const char* path, int64_t* out, int64_t cap) {
    const int fd = ::open(path, O_RDONLY);
    if (fd < 0) return -1;
    struct stat st;
    if (::fstat(fd, &st) != 0) { ::close(fd); return -1; }
    const int64_t size = st.st_size;
    if (size == 0) { ::close(fd); return 0; }
    void* map = ::mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
    ::close(fd);
    if (map == MAP_FAILED) return -1;
    ::madvise(map, size, MADV_SEQUENTIAL);
    const char* base = static_cast<const char*>(map);
    const char* end = base + size;
    int64_t lines = 0, bad = 0;
    const char* p = base;
    while (p < end) {
        const char* nl = static_cast<const char*>(std::memchr(p, '\n', end - p));
        const char* le = nl ? nl : end;
        if (out != nullptr && lines < cap) {
            const int64_t n = parse_line(p, le, nullptr, 0);
            if (n < 0) { bad = -(lines + 2); break; }
            out[2 * lines] = p - base;
            out[2 * lines + 1] = n;
        }
        ++lines;
        p = nl ? nl + 1 : end;
    }
    ::munmap(map, size);
    return bad != 0 ? bad : lines;
}

}  // extern "C"
