// kf_host.cpp -- host-only parts of libkfcount.so: vocabulary, Python-repr-exact .kf formatting,
// synthetic input generators.  No CUDA in this file.
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

// ---- deterministic RNG for the synthetic generators ---------------------------------------------
struct SplitMix {
    uint64_t s;
    explicit SplitMix(uint64_t seed) : s(seed) {}
    inline uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    inline double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    inline uint64_t below(uint64_t n) { return n ? next() % n : 0; }
};

static uint64_t mix_seed(uint64_t seed, uint64_t id, uint64_t salt) {
    SplitMix m(seed * 0x9E3779B97F4A7C15ull + id * 0xD1B54A32D192ED03ull + salt);
    m.next();
    return m.next();
}

struct GenomePlan {
    double gc;
    std::vector<int64_t> contig_len;
    std::vector<std::pair<int64_t, int64_t>> nruns;  // [start, end) in genome coordinates
};

static void plan_genome(uint64_t seed, int64_t id, int64_t n_bases, GenomePlan &P, int max_contigs = 50, int n_runs = 10) {
    SplitMix r(mix_seed(seed, (uint64_t)id, 1));
    P.gc = 0.30 + 0.40 * r.uniform();
    int64_t c = 1 + (int64_t)r.below(50);
    if (c > max_contigs) c = max_contigs;
    if (n_bases < 1000 * c) c = n_bases / 1000 > 0 ? n_bases / 1000 : 1;
    std::vector<double> w((size_t)c);
    double sw = 0;
    for (auto &x : w) { x = r.uniform() + 1e-9; sw += x; }
    int64_t minlen = (n_bases >= 1000 * c) ? 1000 : 0;
    int64_t rem = n_bases - minlen * c, used = 0;
    P.contig_len.assign((size_t)c, 0);
    for (int64_t j = 0; j < c; j++) {
        int64_t extra = (j == c - 1) ? rem - used : (int64_t)((double)rem * (w[(size_t)j] / sw));
        if (extra > rem - used) extra = rem - used;
        P.contig_len[(size_t)j] = minlen + extra;
        used += extra;
    }
    P.nruns.clear();
    if (n_bases > 200) {
        for (int i = 0; i < n_runs; i++) {
            int64_t len = 1 + (int64_t)r.below(100);
            int64_t s = (int64_t)r.below((uint64_t)(n_bases - len));
            P.nruns.push_back({s, s + len});
        }
    }
}

// bases [0,n) of genome (seed,id) into dst (ASCII upper case, N-runs applied)
static void gen_bases(uint64_t seed, int64_t id, const GenomePlan &P, int64_t n, uint8_t *dst) {
    SplitMix r(mix_seed(seed, (uint64_t)id, 2));
    const uint32_t thr = (uint32_t)(P.gc * 32768.0);
    int64_t i = 0;
    while (i < n) {
        uint64_t z = r.next();
        for (int b = 0; b < 4 && i < n; b++, i++) {
            uint32_t v = (uint32_t)(z >> (16 * b)) & 0xFFFFu;
            bool gcb = (v & 0x7FFFu) < thr;
            bool hi = (v >> 15) != 0;
            dst[i] = gcb ? (hi ? 'G' : 'C') : (hi ? 'T' : 'A');
        }
    }
    for (auto &nr : P.nruns)
        for (int64_t p = nr.first; p < nr.second && p < n; p++) dst[p] = 'N';
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

int64_t kf_synth_fasta(uint64_t seed, int64_t genome_id, int64_t n_bases, int line_width, uint8_t *out,
                       size_t out_len) {
    return kf_synth_fasta_ex(seed, genome_id, n_bases, line_width, 50, 10, out, out_len);
}

int64_t kf_synth_fasta_ex(uint64_t seed, int64_t genome_id, int64_t n_bases, int line_width, int max_contigs,
                          int n_runs, uint8_t *out, size_t out_len) {
    if (n_bases < 0 || line_width < 1 || max_contigs < 1 || n_runs < 0) return KF_ERR_ARG;
    kf::GenomePlan P;
    kf::plan_genome(seed, genome_id, n_bases, P, max_contigs, n_runs);
    // size
    int64_t total = 0;
    std::vector<std::string> hdr(P.contig_len.size());
    for (size_t j = 0; j < P.contig_len.size(); j++) {
        char h[64];
        snprintf(h, sizeof h, ">g%05lld_c%d synthetic\n", (long long)genome_id, (int)j);
        hdr[j] = h;
        int64_t L = P.contig_len[j];
        total += (int64_t)hdr[j].size() + L + (L + line_width - 1) / line_width;
    }
    if (!out) return total;
    if ((int64_t)out_len < total) return KF_ERR_ARG;
    std::vector<uint8_t> bases((size_t)n_bases);
    kf::gen_bases(seed, genome_id, P, n_bases, bases.data());
    uint8_t *p = out;
    int64_t g = 0;
    for (size_t j = 0; j < P.contig_len.size(); j++) {
        memcpy(p, hdr[j].data(), hdr[j].size()); p += hdr[j].size();
        int64_t L = P.contig_len[j];
        for (int64_t o = 0; o < L; o += line_width) {
            int64_t m = (L - o < line_width) ? L - o : line_width;
            memcpy(p, bases.data() + g + o, (size_t)m); p += m;
            *p++ = '\n';
        }
        g += L;
    }
    return (int64_t)(p - out);
}

int64_t kf_synth_fastq(uint64_t seed, int64_t sample_id, int64_t genome_len, int64_t n_reads,
                       int read_len, uint8_t *out, size_t out_len) {
    if (genome_len < read_len || n_reads < 0 || read_len < 1) return KF_ERR_ARG;
    // header "@g%05lld.%lld/1\n" has a variable width: compute exactly
    auto hdr_len = [&](int64_t r) {
        char h[64];
        return (int64_t)snprintf(h, sizeof h, "@g%05lld.%lld/1\n", (long long)sample_id, (long long)r);
    };
    int64_t total = 0;
    {
        // digits of r change at powers of ten
        int64_t r = 0;
        while (r < n_reads) {
            int64_t hl = hdr_len(r);
            int64_t next = 10;
            while (next <= r) next *= 10;
            int64_t hi = next < n_reads ? next : n_reads;
            total += (hi - r) * (hl + (int64_t)read_len + 1 + 2 + (int64_t)read_len + 1);
            r = hi;
        }
    }
    if (!out) return total;
    if ((int64_t)out_len < total) return KF_ERR_ARG;
    kf::GenomePlan P;
    kf::plan_genome(seed, sample_id, genome_len, P);
    P.nruns.clear();
    std::vector<uint8_t> g((size_t)genome_len);
    kf::gen_bases(seed, sample_id, P, genome_len, g.data());
    kf::SplitMix r(kf::mix_seed(seed, (uint64_t)sample_id, 3));
    uint8_t *p = out;
    for (int64_t i = 0; i < n_reads; i++) {
        p += snprintf((char *)p, 64, "@g%05lld.%lld/1\n", (long long)sample_id, (long long)i);
        uint64_t z = r.next();
        int64_t s = (int64_t)(z % (uint64_t)(genome_len - read_len + 1));
        bool rc = (z >> 63) != 0;
        uint8_t *seq = p;
        if (!rc) memcpy(seq, g.data() + s, (size_t)read_len);
        else
            for (int j = 0; j < read_len; j++) {
                uint8_t c = g[(size_t)(s + read_len - 1 - j)];
                seq[j] = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : 'N';
            }
        // per-base N with probability ~0.2 % (131/65536)
        for (int j = 0; j < read_len; j += 4) {
            uint64_t q = r.next();
            for (int b = 0; b < 4 && j + b < read_len; b++)
                if (((q >> (16 * b)) & 0xFFFFu) < 131u) seq[j + b] = 'N';
        }
        uint64_t y = r.next();
        if (y % 100 == 0) {  // 1 % of reads carry an N-run of 1..20
            int len = 1 + (int)((y >> 8) % 20);
            int st = (int)((y >> 16) % (uint64_t)read_len);
            for (int j = st; j < st + len && j < read_len; j++) seq[j] = 'N';
        }
        p += read_len;
        *p++ = '\n'; *p++ = '+'; *p++ = '\n';
        for (int j = 0; j < read_len; j += 8) {
            uint64_t q = r.next();
            for (int b = 0; b < 8 && j + b < read_len; b++) p[j + b] = (uint8_t)(33 + ((q >> (8 * b)) & 0xFF) % 42);
        }
        p += read_len;
        *p++ = '\n';
    }
    return (int64_t)(p - out);
}

}  // extern "C"
