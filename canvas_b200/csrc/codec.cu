// Text codecs of the .binned / .cleaned files (reference CanvasCommon/IO.cs:15-52): host code, multi-threaded.
//
// Once the numeric path of a module takes milliseconds, parsing and formatting three million tab-separated lines is
// what a module's wall clock consists of.  These two entry points replace the per-line
// string.Format / Split / Parse loops of CanvasIO.WriteToTextFile and CanvasIO.ReadFromTextFile on the uncompressed
// text; gzip stays with the caller.  No device work, no ctx.
//
//   cg_format_bins  chr \t start \t stop \t count \t gc \n   count as .NET Core 2.0 prints {0:F2} (IO.cs:21): the float's
//                   seven significant decimal digits, then half-up to two decimals; or (four_columns) the merged
//                   pedigree layout of CanvasRunner.cs:895-897 with float.ToString() (general format, 7 digits)
//   cg_parse_bins   the inverse: int.Parse / float.Parse columns (decimal -> double -> float as .NET Core 2.0 does),
//                   chromosome RUN ids in file order
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "clean.cuh"

namespace {

int codec_threads(int64_t n, int requested) {
    int t = requested > 0 ? requested : (int)std::thread::hardware_concurrency();
    if (t < 1) t = 1;
    if (t > 64) t = 64;
    const int64_t by_work = std::max<int64_t>(1, n / 20000);
    return (int)std::min<int64_t>(t, by_work);
}

inline char* put_uint(char* p, unsigned long long v) {
    char tmp[24];
    int k = 0;
    do { tmp[k++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (k) *p++ = tmp[--k];
    return p;
}

inline char* put_int(char* p, long long v) {
    if (v < 0) { *p++ = '-'; return put_uint(p, (unsigned long long)(-(v + 1)) + 1ull); }
    return put_uint(p, (unsigned long long)v);
}

// {0:F2} of a float: the hundredths the .cleaned round trip produces (dotnet_f2_roundtrip), printed as digits
inline char* put_f2(char* p, float v) {
    if (v != v) { memcpy(p, "NaN", 3); return p + 3; }
    if (v - v != 0.0f) {
        if (v < 0) *p++ = '-';
        memcpy(p, "Infinity", 8);
        return p + 8;
    }
    const double r = dotnet_f2_roundtrip(v);  // |r| * 100 is an integer by construction
    const double h = std::nearbyint(std::fabs(r) * 100.0);
    if (h >= 1.8e19) {  // beyond 64-bit hundredths: floats above ~1.8e17 have no fractional digits
        const int len = snprintf(p, 64, "%.0f", (double)v);
        memcpy(p + len, ".00", 3);
        return p + len + 3;
    }
    const unsigned long long hh = (unsigned long long)h;
    if (v < 0 && hh != 0) *p++ = '-';  // -0.00 prints as "0.00" in .NET Core 2.0? it prints "-0.00" only from 3.0 on
    p = put_uint(p, hh / 100);
    *p++ = '.';
    *p++ = (char)('0' + (hh / 10) % 10);
    *p++ = (char)('0' + hh % 10);
    return p;
}

// float.ToString() of .NET Core 2.0: "G7"
inline char* put_g7(char* p, float v) {
    if (v != v) { memcpy(p, "NaN", 3); return p + 3; }
    if (v - v != 0.0f) {
        if (v < 0) *p++ = '-';
        memcpy(p, "Infinity", 8);
        return p + 8;
    }
    char tmp[48];
    snprintf(tmp, sizeof(tmp), "%.7g", (double)v);
    char* e = strchr(tmp, 'e');
    if (!e) {
        const size_t len = strlen(tmp);
        memcpy(p, tmp, len);
        return p + len;
    }
    const size_t ml = (size_t)(e - tmp);
    memcpy(p, tmp, ml);
    p += ml;
    int ex = atoi(e + 1);
    *p++ = 'E';
    *p++ = ex < 0 ? '-' : '+';
    if (ex < 0) ex = -ex;
    if (ex < 10) *p++ = '0';
    return put_uint(p, (unsigned long long)ex);
}

}  // namespace

extern "C" int64_t cg_format_bins(int64_t n, int n_names, const char* const* names, const uint8_t* chrom, const int32_t* start,
                                  const int32_t* stop, const float* count, const uint8_t* gc, int four_columns, char* out,
                                  int64_t cap, int n_threads) {
    if (n < 0 || n_names < 0 || (n > 0 && (!names || !chrom || !start || !stop || !count)) || (!four_columns && n > 0 && !gc)) return -1;
    size_t max_name = 0;
    std::vector<size_t> name_len((size_t)n_names);
    for (int i = 0; i < n_names; i++) { name_len[(size_t)i] = names[i] ? strlen(names[i]) : 0; max_name = std::max(max_name, name_len[(size_t)i]); }
    const int T = codec_threads(n, n_threads);
    const size_t max_line = max_name + 96;  // name + two ints + a float in fixed notation + gc + separators
    struct Part { char* buf = nullptr; size_t len = 0; };
    std::vector<Part> parts((size_t)T);
    std::vector<int> bad((size_t)T, 0);
    auto work = [&](int t) {
        const int64_t lo = n * t / T, hi = n * (t + 1) / T;
        size_t cap = (size_t)(hi - lo) * (max_name + 36) + max_line;  // typical line; grown when a thread runs ahead of it
        char* buf = (char*)malloc(cap);
        if (!buf) { bad[(size_t)t] = 2; return; }
        char* p = buf;
        for (int64_t i = lo; i < hi; i++) {
            if ((size_t)(buf + cap - p) < max_line) {
                const size_t used = (size_t)(p - buf);
                cap = cap + cap / 2 + max_line;
                char* nb = (char*)realloc(buf, cap);
                if (!nb) { free(buf); bad[(size_t)t] = 2; return; }
                buf = nb;
                p = buf + used;
            }
            const int c = chrom[i];
            if (c >= n_names) { bad[(size_t)t] = 1; continue; }
            memcpy(p, names[c], name_len[(size_t)c]);
            p += name_len[(size_t)c];
            *p++ = '\t';
            p = put_int(p, start[i]);
            *p++ = '\t';
            p = put_int(p, stop[i]);
            *p++ = '\t';
            if (four_columns) p = put_g7(p, count[i]);
            else {
                p = put_f2(p, count[i]);
                *p++ = '\t';
                p = put_uint(p, gc[i]);
            }
            *p++ = '\n';
        }
        parts[(size_t)t].buf = buf;
        parts[(size_t)t].len = (size_t)(p - buf);
    };
    std::vector<std::thread> th;
    for (int t = 1; t < T; t++) th.emplace_back(work, t);
    work(0);
    for (auto& x : th) x.join();
    int64_t total = 0;
    bool failed = false;
    for (int t = 0; t < T; t++) { failed = failed || bad[(size_t)t]; total += (int64_t)parts[(size_t)t].len; }
    if (!failed && out && cap >= total) {
        // the pieces are copied to their places side by side as well
        std::vector<size_t> at((size_t)T, 0);
        for (int t = 1; t < T; t++) at[(size_t)t] = at[(size_t)t - 1] + parts[(size_t)t - 1].len;
        std::vector<std::thread> cp;
        for (int t = 1; t < T; t++) cp.emplace_back([&, t] { memcpy(out + at[(size_t)t], parts[(size_t)t].buf, parts[(size_t)t].len); });
        memcpy(out, parts[0].buf, parts[0].len);
        for (auto& x : cp) x.join();
    }
    for (int t = 0; t < T; t++) free(parts[(size_t)t].buf);
    return failed ? -1 : total;  // out == NULL or cap < total: size query
}

// Returns the number of rows (> max_rows: buffers too small, nothing complete), or < 0: -1 bad argument, -2 malformed
// line (fewer than four columns or a number that does not parse), -3 more than 256 chromosome runs / names do not fit.
extern "C" int64_t cg_parse_bins(const char* text, int64_t len, int64_t max_rows, uint8_t* chrom, int32_t* start, int32_t* stop,
                                 float* count, uint8_t* gc, int* n_names, char* names, int64_t names_cap, int n_threads) {
    if (len < 0 || (len > 0 && !text) || !n_names) return -1;
    *n_names = 0;
    const int T = codec_threads(len / 24 + 1, n_threads);
    // chunk boundaries at line starts
    std::vector<int64_t> cut((size_t)T + 1, len);
    cut[0] = 0;
    for (int t = 1; t < T; t++) {
        int64_t p = len * t / T;
        while (p < len && text[p - 1] != '\n') p++;
        cut[(size_t)t] = std::max(p, cut[(size_t)t - 1]);
    }
    struct Row { int64_t name_off; int name_len; int32_t a, b; float v; uint8_t g; };
    std::vector<std::vector<Row>> rows((size_t)T);
    std::vector<int> err((size_t)T, 0);
    auto work = [&](int t) {
        std::vector<Row>& out = rows[(size_t)t];
        out.reserve((size_t)((cut[(size_t)t + 1] - cut[(size_t)t]) / 20 + 16));
        const char* p = text + cut[(size_t)t];
        const char* end = text + cut[(size_t)t + 1];
        char num[64];
        while (p < end) {
            const char* eol = (const char*)memchr(p, '\n', (size_t)(end - p));
            if (!eol) eol = end;
            const char* le = eol;
            if (le > p && le[-1] == '\r') le--;
            if (le > p) {
                const char* f[5];
                int flen[5];
                int nf = 0;
                const char* q = p;
                while (nf < 5) {
                    const char* tab = (const char*)memchr(q, '\t', (size_t)(le - q));
                    f[nf] = q;
                    flen[nf] = (int)((tab ? tab : le) - q);
                    nf++;
                    if (!tab) break;
                    q = tab + 1;
                }
                if (nf < 4) { err[(size_t)t] = 2; return; }
                Row r;
                r.name_off = f[0] - text;
                r.name_len = flen[0];
                auto to_num = [&](int k) -> bool {
                    if (flen[k] <= 0 || flen[k] >= (int)sizeof(num)) return false;
                    memcpy(num, f[k], (size_t)flen[k]);
                    num[flen[k]] = 0;
                    return true;
                };
                char* stop_at = nullptr;
                if (!to_num(1)) { err[(size_t)t] = 2; return; }
                long long a = strtoll(num, &stop_at, 10);
                if (*stop_at) { err[(size_t)t] = 2; return; }
                if (!to_num(2)) { err[(size_t)t] = 2; return; }
                long long b = strtoll(num, &stop_at, 10);
                if (*stop_at) { err[(size_t)t] = 2; return; }
                if (!to_num(3)) { err[(size_t)t] = 2; return; }
                // .NET Core 2.0 float.Parse: the decimal is converted to double and then narrowed (Number.NumberToSingle)
                const float v = (float)strtod(num, &stop_at);
                if (*stop_at) { err[(size_t)t] = 2; return; }
                long long g = 0;
                if (nf >= 5) {
                    if (!to_num(4)) { err[(size_t)t] = 2; return; }
                    g = strtoll(num, &stop_at, 10);
                    if (*stop_at) { err[(size_t)t] = 2; return; }
                }
                if (a < INT32_MIN || a > INT32_MAX || b < INT32_MIN || b > INT32_MAX || g < 0 || g > 255) { err[(size_t)t] = 2; return; }
                r.a = (int32_t)a; r.b = (int32_t)b; r.v = v; r.g = (uint8_t)g;
                out.push_back(r);
            }
            p = eol + 1;
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < T; t++) th.emplace_back(work, t);
    work(0);
    for (auto& x : th) x.join();
    int64_t total = 0;
    for (int t = 0; t < T; t++) { if (err[(size_t)t]) return -(int64_t)err[(size_t)t]; total += (int64_t)rows[(size_t)t].size(); }
    if (total > max_rows || (total > 0 && (!chrom || !start || !stop || !count))) return total > max_rows ? total : -1;
    // chromosome runs in file order (a name that reappears later starts a new run, as CanvasClean's run-based loops see it)
    int64_t k = 0, names_used = 0;
    int runs = 0;
    const char* prev = nullptr;
    int prev_len = -1;
    for (int t = 0; t < T; t++)
        for (const Row& r : rows[(size_t)t]) {
            const char* nm = text + r.name_off;
            if (prev_len != r.name_len || memcmp(prev, nm, (size_t)r.name_len) != 0) {
                if (runs >= 256 || !names || names_used + r.name_len + 1 > names_cap) return -3;
                memcpy(names + names_used, nm, (size_t)r.name_len);
                names[names_used + r.name_len] = 0;
                names_used += r.name_len + 1;
                runs++;
                prev = nm;
                prev_len = r.name_len;
            }
            chrom[k] = (uint8_t)(runs - 1);
            start[k] = r.a; stop[k] = r.b; count[k] = r.v;
            if (gc) gc[k] = r.g;
            k++;
        }
    *n_names = runs;
    return total;
}
