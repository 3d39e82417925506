// kf_host.cpp -- host-only parts of libkfcount.so: vocabulary, Python-repr-exact .kf formatting and parsing,
// chunked-mode text preparation.  No CUDA in this file.
//
// Reference behaviour restated (paths in the kf2vec checkout):
//   vocabulary order            kf2vec/main.py:278-296 + kf2vec/data/test_kmers_7_sorted etc.
//   astype(str) + join + write  kf2vec/main.py:344-357
#include "kfcount.h"

#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <thread>
#include <atomic>

namespace kf {

uint32_t revcomp_std(uint32_t x, int k) {
    uint32_t o = 0;
    for (int i = 0; i < k; i++) { o = (o << 2) | (3u - (x & 3u)); x >>= 2; }
    return o;
}

// canonical codes in sorted order (A0 C1 G2 T3)
void canonical_codes(int k, std::vector<uint32_t> &out) {
    out.clear();
    const uint64_t nb = 1ull << (2 * k);
    out.reserve((size_t)kf_vocab_size(k));
    for (uint64_t x = 0; x < nb; x++) {
        uint32_t r = revcomp_std((uint32_t)x, k);
        if ((uint32_t)x <= r) out.push_back((uint32_t)x);
    }
}

// ---- Python repr(float) ------------------------------------------------------------------------
// Shortest round-trip digits (std::to_chars), laid out by CPython's float_repr_style='short' rule:
// fixed notation when -4 <= exp10 < 16, otherwise d[.ddd]e+XX with at least two exponent digits.
int format_repr(double v, char *out) {
    if (std::isnan(v)) { memcpy(out, "nan", 3); return 3; }
    if (std::isinf(v)) {
        if (v < 0) { memcpy(out, "-inf", 4); return 4; }
        memcpy(out, "inf", 3); return 3;
    }
    char *p = out;
    if (std::signbit(v)) { *p++ = '-'; v = -v; }
    if (v == 0.0) { memcpy(p, "0.0", 3); return (int)(p - out) + 3; }
    char buf[48];
    auto res = std::to_chars(buf, buf + sizeof buf, v, std::chars_format::scientific);
    // buf = d[.ddd]e[+-]XX
    char digits[32];
    int nd = 0;
    char *q = buf;
    while (q < res.ptr && *q != 'e') { if (*q != '.') digits[nd++] = *q; q++; }
    int e = 0;
    {
        q++;  // 'e'
        int sign = 1;
        if (*q == '-') { sign = -1; q++; } else if (*q == '+') q++;
        while (q < res.ptr) { e = e * 10 + (*q - '0'); q++; }
        e *= sign;
    }
    if (e >= -4 && e < 16) {
        if (e >= 0) {
            for (int i = 0; i <= e; i++) *p++ = (i < nd) ? digits[i] : '0';
            *p++ = '.';
            if (nd > e + 1) { for (int i = e + 1; i < nd; i++) *p++ = digits[i]; }
            else *p++ = '0';
        } else {
            *p++ = '0'; *p++ = '.';
            for (int i = 0; i < -e - 1; i++) *p++ = '0';
            for (int i = 0; i < nd; i++) *p++ = digits[i];
        }
    } else {
        *p++ = digits[0];
        if (nd > 1) { *p++ = '.'; for (int i = 1; i < nd; i++) *p++ = digits[i]; }
        *p++ = 'e';
        *p++ = (e < 0) ? '-' : '+';
        int ae = e < 0 ? -e : e;
        char eb[8]; int ne = 0;
        while (ae > 0) { eb[ne++] = (char)('0' + ae % 10); ae /= 10; }
        while (ne < 2) eb[ne++] = '0';
        while (ne > 0) *p++ = eb[--ne];
    }
    return (int)(p - out);
}

}  // namespace kf

extern "C" {

int kf_abi_version(void) { return KF_ABI_VERSION; }

const char *kf_strerror(int code) {
    switch (code) {
        case KF_OK: return "ok";
        case KF_ERR_ARG: return "bad argument";
        case KF_ERR_NO_DEVICE: return "no usable CUDA device (sm_100 required; call kf_init first)";
        case KF_ERR_CUDA: return "CUDA runtime error (see kf_last_cuda_error)";
        case KF_ERR_IO: return "file I/O error";
        case KF_ERR_FORMAT: return "unsupported format: first byte is neither '>' nor '@'";
        case KF_ERR_FASTQ: return "FASTQ is not in 4-line layout";
        case KF_ERR_NOMEM: return "out of memory";
        case KF_ERR_LAYOUT: return "device arena violates the layout contract";
        case KF_ERR_EMPTY: return "empty input";
        case KF_ERR_UNSUPPORTED: return "input format not supported by this entry point";
        default: return "unknown error";
    }
}

int64_t kf_vocab_size(int k) {
    if (k < KF_MIN_K || k > 15) return KF_ERR_ARG;
    const int64_t nb = 1ll << (2 * k);
    return (nb + ((k % 2 == 0) ? (1ll << k) : 0)) / 2;
}

int kf_vocab_codes(int k, uint32_t *out, size_t n_out) {
    if (k < KF_MIN_K || k > KF_MAX_K || !out) return KF_ERR_ARG;
    std::vector<uint32_t> c;
    kf::canonical_codes(k, c);
    if (n_out < c.size()) return KF_ERR_ARG;
    memcpy(out, c.data(), c.size() * sizeof(uint32_t));
    return KF_OK;
}

int kf_vocab(int k, char *out, size_t out_len) {
    if (k < KF_MIN_K || k > KF_MAX_K || !out) return KF_ERR_ARG;
    std::vector<uint32_t> c;
    kf::canonical_codes(k, c);
    if (out_len < c.size() * (size_t)(k + 1)) return KF_ERR_ARG;
    char *p = out;
    for (uint32_t x : c) {
        for (int i = 0; i < k; i++) *p++ = "ACGT"[(x >> (2 * (k - 1 - i))) & 3u];
        *p++ = '\n';
    }
    return KF_OK;
}

int64_t kf_format_row(const char *sample, const double *row, int64_t V, int int_mode, char *out,
                      size_t out_len) {
    if (!sample || !row || !out || V < 0) return KF_ERR_ARG;
    const size_t sl = strlen(sample);
    // worst case per value: sign + 17 digits + '.' + "e-308" + ',' < 32
    if (out_len < sl + 2 + (size_t)V * 32) return KF_ERR_ARG;
    char *p = out;
    memcpy(p, sample, sl); p += sl;
    for (int64_t i = 0; i < V; i++) {
        *p++ = ',';
        if (int_mode) {
            auto r = std::to_chars(p, p + 24, (long long)row[i]);
            p = r.ptr;
        } else {
            const double v = row[i];
            if (v >= 0.0 && v < 1e15 && v == (double)(long long)v) {   // raw counts: "16.0" without the shortest-digits search
                auto r = std::to_chars(p, p + 24, (long long)v);
                p = r.ptr;
                *p++ = '.';
                *p++ = '0';
            } else {
                p += kf::format_repr(v, p);
            }
        }
    }
    *p++ = '\n';
    *p = 0;
    return (int64_t)(p - out);
}

// Several rows into one file with one open: the chunked-genome output (main.py:895-915 concatenates one row per chunk).
// labels: n NUL-terminated strings back to back; int_modes[i] as in kf_format_row.
int kf_write_kf_rows(const char *out_path, const char *labels, const double *rows, int64_t n, int64_t V, const uint8_t *int_modes,
                     int append) {
    if (!out_path || !labels || !rows || n < 0 || V < 0) return KF_ERR_ARG;
    FILE *f = fopen(out_path, append ? "ab" : "wb");
    if (!f) return KF_ERR_IO;
    std::vector<const char *> lab((size_t)n);
    {
        const char *p = labels;
        for (int64_t i = 0; i < n; i++) { lab[(size_t)i] = p; p += strlen(p) + 1; }
    }
    // rows are formatted by several threads in blocks of consecutive rows (a row is ~8,192 numbers), each block into its
    // own buffer; the blocks are written in order
    const int64_t BLOCK = 16;
    const int64_t nblocks = (n + BLOCK - 1) / BLOCK;
    const int T = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(nblocks, 16), (int64_t)std::thread::hardware_concurrency()));
    std::vector<std::vector<char>> out((size_t)nblocks);
    std::atomic<int64_t> next(0);
    std::atomic<int> rc(KF_OK);
    auto work = [&]() {
        for (;;) {
            const int64_t b = next.fetch_add(1);
            if (b >= nblocks) return;
            std::vector<char> &buf = out[(size_t)b];
            size_t used = 0;
            for (int64_t i = b * BLOCK; i < std::min(n, (b + 1) * BLOCK); i++) {
                const size_t need = strlen(lab[(size_t)i]) + 2 + (size_t)V * 32 + 1;
                if (buf.size() < used + need) buf.resize(std::max(buf.size() * 2, used + need));
                const int64_t m = kf_format_row(lab[(size_t)i], rows + (size_t)i * (size_t)V, V, int_modes ? int_modes[i] : 0, buf.data() + used, need);
                if (m < 0) { rc.store((int)m); return; }
                used += (size_t)m;
            }
            buf.resize(used);
        }
    };
    if (T <= 1) work();
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < T; t++) th.emplace_back(work);
        for (auto &t : th) t.join();
    }
    int r = rc.load();
    for (int64_t b = 0; b < nblocks && r == KF_OK; b++)
        if (!out[(size_t)b].empty() && fwrite(out[(size_t)b].data(), 1, out[(size_t)b].size(), f) != out[(size_t)b].size()) r = KF_ERR_IO;
    if (fclose(f) != 0 && r == KF_OK) r = KF_ERR_IO;
    return r;
}

// ---- chunked-genome text preparation: main.py:730-753 ------------------------------------------------------------
// One pass over a FASTA buffer doing what the reference shells out for: `seqtk seq -l 0` (linearise, :732), awk
// gsub(/[N|n]+/,"N") on sequence lines (:740 -- the class holds 'N', '|' and 'n'), `seqkit seq -g -m min_len` (:753:
// remove the gap characters '-', '.', ' ', then drop records shorter than min_len).  The kept sequences are written back
// to back into seq_out; record r is seq_out[seq_off[r] .. +seq_len[r]) and its header text is data[id_off[r] .. +id_len[r])
// (the whole header line without '>'; the contig id is its first white-space separated token).
// Returns the number of kept records (only the first max_records are stored), or a negative error.
int64_t kf_linearise_fasta(const uint8_t *data, size_t len, uint64_t min_len, uint8_t *seq_out, size_t seq_cap, uint64_t *seq_off,
                           uint64_t *seq_len, uint64_t *id_off, uint32_t *id_len, int64_t max_records) {
    if (!data || !seq_out) return KF_ERR_ARG;
    if (len == 0 || data[0] != '>') return 0;
    size_t p = 0, w = 0;
    int64_t kept = 0;
    while (p < len) {
        // header line (p is at a '>')
        const size_t h0 = p + 1;
        const uint8_t *nl = (const uint8_t *)memchr(data + p, '\n', len - p);
        size_t h1 = nl ? (size_t)(nl - data) : len;
        p = nl ? h1 + 1 : len;
        if (h1 > h0 && data[h1 - 1] == '\r') h1--;
        const size_t start = w;
        bool in_run = false;   // inside a run of N / n / |
        // sequence lines up to the next line that begins with '>'
        while (p < len && data[p] != '>') {
            const uint8_t *e = (const uint8_t *)memchr(data + p, '\n', len - p);
            const size_t le = e ? (size_t)(e - data) : len;
            if (w + (le - p) > seq_cap) return KF_ERR_ARG;
            // most lines hold none of the bytes treated specially below: one vectorisable scan, then a plain copy
            unsigned special = 0;
            for (size_t i = p; i < le; i++) {
                const uint8_t c = data[i];
                special |= (unsigned)(c == 'N') | (unsigned)(c == 'n') | (unsigned)(c == '|') | (unsigned)(c == '\r') | (unsigned)(c == '-') |
                           (unsigned)(c == '.') | (unsigned)(c == ' ');
            }
            if (!special) {
                if (le > p) { memcpy(seq_out + w, data + p, le - p); w += le - p; in_run = false; }
                p = e ? le + 1 : len;
                continue;
            }
            for (size_t i = p; i < le; i++) {
                const uint8_t c = data[i];
                if (c == 'N' || c == 'n' || c == '|') { if (!in_run) seq_out[w++] = 'N'; in_run = true; continue; }
                if (c == '\r') continue;   // line ends are dropped before the runs are collapsed (a run may span lines)
                in_run = false;
                if (c == '-' || c == '.' || c == ' ') continue;   // gaps go after the collapse: "N-N" stays two Ns
                seq_out[w++] = c;
            }
            p = e ? le + 1 : len;
        }
        if (w - start >= min_len) {
            if (kept < max_records) {
                if (seq_off) seq_off[kept] = start;
                if (seq_len) seq_len[kept] = w - start;
                if (id_off) id_off[kept] = h0;
                if (id_len) id_len[kept] = (uint32_t)(h1 - h0);
            }
            kept++;
        } else {
            w = start;   // dropped: reuse the space
        }
    }
    return kept;
}

int kf_write_kf(const char *out_path, const char *sample, const double *row, int64_t V, int int_mode,
                int append) {
    if (!out_path || !sample || !row) return KF_ERR_ARG;
    std::vector<char> buf(strlen(sample) + 2 + (size_t)V * 32 + 1);
    int64_t n = kf_format_row(sample, row, V, int_mode, buf.data(), buf.size());
    if (n < 0) return (int)n;
    FILE *f = fopen(out_path, append ? "ab" : "wb");
    if (!f) return KF_ERR_IO;
    size_t w = fwrite(buf.data(), 1, (size_t)n, f);
    int rc = fclose(f);
    return (w == (size_t)n && rc == 0) ? KF_OK : KF_ERR_IO;
}

// ---- .kf reader (utils.py:436-437 my_read_csv; classify.py:102-114; query.py:148-158) ----------------------------
// Parses the rows "label,v1,...,vV\n" of a .kf text buffer.  Values are converted with std::from_chars (correctly
// rounded, so repr()-formatted numbers round-trip bit-exactly); "nan" parses to NaN; integers ("5") are accepted.
// out rows are written as double [row][V]; if feat_out is given it receives float(value * 1e4), the tensor the
// trainers build (train_classifier_model.py:149,323).  label_off/label_len locate each row's label in text.
// Returns the number of rows parsed (may exceed max_rows: then only the first max_rows were stored), or a negative
// error: KF_ERR_FORMAT when a row has other than V values or a value does not parse.
int64_t kf_parse_kf(const char *text, size_t len, int64_t V, int64_t max_rows, double *out, float *feat_out,
                    int64_t *label_off, int32_t *label_len) {
    if (!text || V < 0 || max_rows < 0) return KF_ERR_ARG;
    const char *p = text, *end = text + len;
    int64_t row = 0;
    while (p < end) {
        const char *eol = (const char *)memchr(p, '\n', (size_t)(end - p));
        if (!eol) eol = end;
        const char *le = eol;
        if (le > p && le[-1] == '\r') le--;
        if (le == p) { p = eol + 1; continue; }   // blank line
        const char *c = (const char *)memchr(p, ',', (size_t)(le - p));
        const bool store = row < max_rows;
        if (store && label_off) label_off[row] = (int64_t)(p - text);
        if (store && label_len) label_len[row] = (int32_t)((c ? c : le) - p);
        int64_t nv = 0;
        const char *q = c ? c + 1 : le;
        while (c && q <= le) {
            const char *ve = (const char *)memchr(q, ',', (size_t)(le - q));
            if (!ve) ve = le;
            double v;
            if (ve - q == 3 && (q[0] == 'n' || q[0] == 'N') && (q[1] == 'a' || q[1] == 'A') && (q[2] == 'n' || q[2] == 'N')) v = NAN;
            else {
                auto r = std::from_chars(q, ve, v);
                if (r.ec != std::errc() || r.ptr != ve) return KF_ERR_FORMAT;
            }
            if (nv >= V) return KF_ERR_FORMAT;
            if (store && out) out[row * V + nv] = v;
            if (store && feat_out) feat_out[row * V + nv] = (float)(v * 1e4);
            nv++;
            if (ve == le) break;
            q = ve + 1;
        }
        if (nv != V) return KF_ERR_FORMAT;
        row++;
        p = eol + 1;
    }
    return row;
}

}  // extern "C"
