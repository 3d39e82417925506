// kf_synth.cpp -- deterministic synthetic inputs for bench.py, the tests and tools/ (SURVEY.md section 8d configs 2 and 4).
// Test/bench infrastructure: built into tools/libkfsynth.so, NOT part of the product library libkfcount.so.
// The genomes come from a private SplitMix64 stream, not from numpy.random.default_rng(1000 + i) as SURVEY.md 8d sketched
// (1,000 x 5 Mbp have to be generated in seconds on the GPU box): same distributional recipe -- GC ~ U(0.30, 0.70), 1..50
// contigs of at least 1 kbp, 10 N-runs of 1..100, upper case, fixed-width lines, LF.
#include "kf_synth.h"

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#define KF_ERR_ARG (-1)

namespace kf {

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

int64_t kf_synth_fasta(uint64_t seed, int64_t genome_id, int64_t n_bases, int line_width, uint8_t *out,
                       size_t out_len) {
    return kf_synth_fasta_ex(seed, genome_id, n_bases, line_width, 50, 10, out, out_len);
}

int64_t kf_synth_fasta_ex(uint64_t seed, int64_t genome_id, int64_t n_bases, int line_width, int max_contigs,
                          int n_runs, uint8_t *out, size_t out_len) {
    if (n_bases < 0 || line_width < 0 || max_contigs < 1 || n_runs < 0) return KF_ERR_ARG;
    const int64_t lw = line_width ? (int64_t)line_width : ((int64_t)1 << 60);   // 0: unwrapped, one line per contig
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
        total += (int64_t)hdr[j].size() + L + (L + lw - 1) / lw;
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
        for (int64_t o = 0; o < L; o += lw) {
            int64_t m = (L - o < lw) ? L - o : lw;
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
