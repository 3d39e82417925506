// kf_kernels.cuh -- device code of the k-mer frequency engine (sm_100a).
//
// Replaces `jellyfish count -m k -C` + `jellyfish dump -c` (kf2vec/main.py:308-319) and the pandas
// merge / pseudocount / normalise (main.py:327-342).
//
// Symbol model (Jellyfish's view of a sequence file, restated in oracle/kf_oracle.py):
//   BASE(code)  A/C/G/T in either case
//   SKIP        '\n' inside a FASTA sequence line: removed, does not break the window
//   BREAK       anything else: N, IUPAC, '\r', NUL padding, every header byte, record boundaries
// A k-mer occurrence is k consecutive BASE symbols with only SKIPs between them.  A k-mer is OWNED by
// the 16-byte lane that holds the byte of its first base, so every occurrence is counted exactly once
// no matter how the arena is cut into tiles, warp ranges, chunks and lanes.
//
// Alphabet: bases are coded (c >> 1) & 3, i.e. A0 C1 T2 G3 ("gray"); forward k-mers are counted in
// that alphabet and mapped to the reference's sorted canonical order (A0 C1 G2 T3) by the fold kernel,
// which also adds the reverse-complement bin: count(min(m, rc m)) = fwd[m] + fwd[rc m].
#pragma once
#ifdef KF_EMU
#include "cuda_emu.h"   // tests/emu: host emulation used by the CPU-side kernel-logic tests only
#else
#include <cuda_runtime.h>
#define KF_DYN_SMEM(type, name) extern __shared__ type name[]
#define KF_NOINLINE __noinline__
#endif
#include <stdint.h>

namespace kf {

constexpr int CHUNK = 512;            // bytes per warp-load (32 lanes x 16 B)
constexpr unsigned FULL = 0xFFFFFFFFu;

struct Tile {
    uint32_t first_chunk;  // arena chunk index
    uint32_t n_chunks;
    uint32_t file;         // file index in the batch
    uint32_t file_chunk0;  // first chunk of that file
};

// ------------------------------------------------------------------------------------------------
// 16-byte decode: 2-bit packing, validity, single-newline compaction
// ------------------------------------------------------------------------------------------------
struct Lane {
    uint32_t bits;     // up to 16 bases, 2 bits each, first base in bits 31:30, zero-filled tail
    uint32_t n;        // 16, or 15 when one '\n' was removed
    uint32_t dirty;    // 1: holds a BREAK byte or more than one '\n' -> byte walker
};

// V == 0 for a byte  <=>  the byte is one of ACGTacgt.  Checked bits: 7,6,4,3,0 (bit 5 = case and
// bits 2:1 = code are free); T/t is the only base with bit4=1/bit0=0, recognised by code == 2.
__device__ __forceinline__ void decode_word(uint32_t w, uint32_t &pk, uint32_t &V) {
    const uint32_t s1 = w >> 1;
    const uint32_t t = s1 & 0x03030303u;                 // gray codes, one per byte
    const uint32_t e = (w >> 2) & ~s1 & 0x01010101u;     // 1 where code == 2 (T)
    const uint32_t ex = e * 0x0Fu + 0x41414141u;         // expected (w & 0xD9): 0x41 or 0x50
    V = (w & 0xD9D9D9D9u) ^ ex;
    pk = t * 0x40100401u;                                // byte 3 = c0<<6 | c1<<4 | c2<<2 | c3
}

__device__ __forceinline__ Lane decode16(const uint4 w) {
    uint32_t p0, p1, p2, p3, V0, V1, V2, V3;
    decode_word(w.x, p0, V0);
    decode_word(w.y, p1, V1);
    decode_word(w.z, p2, V2);
    decode_word(w.w, p3, V3);
    const uint32_t r1 = __byte_perm(p3, p2, 0x0073);     // [.., .., p2.b3, p3.b3]
    const uint32_t r2 = __byte_perm(p1, p0, 0x0073);     // [.., .., p0.b3, p1.b3]
    uint32_t bits = __byte_perm(r1, r2, 0x5410);         // [p0.b3, p1.b3, p2.b3, p3.b3]
    Lane L;
    L.n = 16;
    L.dirty = 0;
    const uint32_t anyV = V0 | V1 | V2 | V3;
    if (anyV) {
        uint32_t wi, Vw, ww, rest;
        if (V0)      { wi = 0; Vw = V0; ww = w.x; rest = V1 | V2 | V3; }
        else if (V1) { wi = 1; Vw = V1; ww = w.y; rest = V2 | V3; }
        else if (V2) { wi = 2; Vw = V2; ww = w.z; rest = V3; }
        else         { wi = 3; Vw = V3; ww = w.w; rest = 0; }
        const uint32_t sh = (uint32_t)(__ffs((int)Vw) - 1) & ~7u;   // bit offset of the first offending byte
        const bool is_nl = ((ww >> sh) & 0xFFu) == 0x0Au;
        rest |= Vw & ~(0xFFu << sh);
        if (is_nl && rest == 0) {
            const uint32_t p = wi * 4 + (sh >> 3);
            const uint32_t m = 0xFFFFFFFFu >> (2 * p);             // fields p..15
            bits = (bits & ~m) | ((bits << 2) & m);                // delete field p
            L.n = 15;
        } else {
            L.dirty = 1;
        }
    }
    L.bits = bits;
    return L;
}

__device__ __forceinline__ uint32_t byte_of(const uint4 &w, int i) {
    const uint32_t x = (i < 4) ? w.x : (i < 8) ? w.y : (i < 12) ? w.z : w.w;
    return (x >> (8 * (i & 3))) & 0xFFu;
}

__device__ __forceinline__ bool is_base(uint32_t c) {
    const uint32_t e = (c >> 2) & ~(c >> 1) & 1u;
    return ((c & 0xD9u) ^ (0x41u + e * 0x0Fu)) == 0;
}

// ------------------------------------------------------------------------------------------------
// FASTA line state.  State at a byte position = (hdr: inside a '>' header line, ls: at a line start)
// ------------------------------------------------------------------------------------------------
// State at the first byte of every lane of one chunk, given the state at the chunk's first byte.
// Returns in_hdr | in_ls << 1 | out_hdr << 2 (state at this lane's first byte; state after the chunk).
__device__ KF_NOINLINE uint32_t fasta_resolve(const uint4 w, bool carry_hdr, bool carry_ls, int lane) {
    bool in_hdr, in_ls, out_hdr;
    bool has_nl = false;
    int last = -1;
#pragma unroll
    for (int i = 0; i < 16; i++)
        if (byte_of(w, i) == 0x0Au) { has_nl = true; last = i; }
    bool after_gt = false;
#pragma unroll
    for (int i = 1; i < 16; i++)
        if (last == i - 1 && byte_of(w, i) == (uint32_t)'>') after_gt = true;
    const bool first_gt = byte_of(w, 0) == (uint32_t)'>';
    const bool o_hdr = has_nl && last < 15 && after_gt;   // state after this lane, if it holds a '\n'
    const bool o_ls = has_nl && last == 15;
    const unsigned B_nl = __ballot_sync(FULL, has_nl);
    const unsigned B_oh = __ballot_sync(FULL, o_hdr);
    const unsigned B_ols = __ballot_sync(FULL, o_ls);
    const unsigned B_fg = __ballot_sync(FULL, first_gt);
    auto state_before = [&](int l, bool &h, bool &s) {
        const unsigned prev = (l >= 32) ? B_nl : (B_nl & ((1u << l) - 1u));
        if (prev) {
            const int j = 31 - __clz((int)prev);
            h = (B_oh >> j) & 1u;
            s = (B_ols >> j) & 1u;
            if (j + 1 < l && s) { h = (B_fg >> (j + 1)) & 1u; s = false; }
        } else {
            h = carry_hdr;
            s = carry_ls;
            if (l > 0 && s) { h = B_fg & 1u; s = false; }
        }
    };
    state_before(lane, in_hdr, in_ls);
    bool dummy;
    state_before(32, out_hdr, dummy);
    return (in_hdr ? 1u : 0u) | (in_ls ? 2u : 0u) | (out_hdr ? 4u : 0u);
}

// State at the first byte of chunk c (file starts at chunk file_c0): scan back to the previous '\n'.
// Returns hdr | ls << 1.
__device__ KF_NOINLINE uint32_t fasta_backscan(const uint8_t *__restrict__ arena, uint32_t c, uint32_t file_c0,
                                            int lane) {
    if (c == file_c0) return 2u;
    const uint4 *base = reinterpret_cast<const uint4 *>(arena);
    for (uint32_t b = c; b-- > file_c0;) {
        const uint4 w = __ldg(base + (size_t)b * 32 + lane);
        int last = -1;
#pragma unroll
        for (int i = 0; i < 16; i++)
            if (byte_of(w, i) == 0x0Au) last = i;
        const unsigned B = __ballot_sync(FULL, last >= 0);
        if (B) {
            const int j = 31 - __clz((int)B);
            const int lastj = __shfl_sync(FULL, last, j);
            const uint64_t q = (uint64_t)b * CHUNK + (uint64_t)j * 16 + (uint64_t)lastj;   // last '\n' before chunk c
            if (q + 1 == (uint64_t)c * CHUNK) return 2u;
            return (arena[q + 1] == (uint8_t)'>') ? 1u : 0u;
        }
    }
    return (arena[(uint64_t)file_c0 * CHUNK] == (uint8_t)'>') ? 1u : 0u;   // still on the file's first line
}

// Byte walker for one lane: counts every k-mer whose first base lies in [p0, p0+16).
template <int K, class Emit>
__device__ KF_NOINLINE void fasta_walk_lane(const uint8_t *__restrict__ arena, uint64_t p0, bool in_hdr,
                                                bool at_ls, Emit emit) {
    constexpr uint32_t MASK = (K >= 16) ? 0xFFFFFFFFu : ((1u << (2 * K)) - 1u);
    uint32_t kmer = 0;
    int run = 0, owned = 0;
    uint64_t p = p0;
    const uint64_t own_end = p0 + 16;
    for (;;) {
        const bool own = p < own_end;
        if (!own && (owned == 0 || run - K + 1 >= owned)) break;
        const uint32_t c = arena[p];
        p++;
        if (in_hdr) {
            if (c == 0x0Au) { in_hdr = false; at_ls = true; }
            run = 0; owned = 0;
            continue;
        }
        if (c == 0x0Au) { at_ls = true; continue; }
        if (at_ls && c == (uint32_t)'>') { in_hdr = true; at_ls = false; run = 0; owned = 0; continue; }
        at_ls = false;
        if (!is_base(c)) { run = 0; owned = 0; continue; }
        kmer = ((kmer << 2) | ((c >> 1) & 3u)) & MASK;
        run++;
        if (own) owned++;
        if (run >= K && run - K < owned) emit(kmer);
    }
}

// ------------------------------------------------------------------------------------------------
// One warp over the chunk range [c0, c1) of a FASTA file
// ------------------------------------------------------------------------------------------------
// Byte offset (4 * k-mer) of the k-mer starting at base j of the 32-base window hi:lo (first base in hi
// bits 31:30).  j is a compile-time constant after unrolling: one shift or funnel shift plus one mask.
template <int K>
__device__ __forceinline__ uint32_t kmer_off_at(uint32_t hi, uint32_t lo, int j) {
    constexpr uint32_t MASK4 = ((1u << (2 * K)) - 1u) << 2;
    const int r = 62 - 2 * K - 2 * j;   // >= 8 for K <= 12, j <= 15
    return ((r >= 32) ? (hi >> (r - 32)) : __funnelshift_r(lo, hi, r)) & MASK4;
}

// PF = 512-byte chunks kept in flight per warp (register ring; the loop is unrolled PF times so the ring
// rotates at compile time).  Sinks take the byte offset 4*kmer.
template <int K, bool FORCE_WALKER, int PF, class Sink>
__device__ __forceinline__ void fasta_process_range(const uint8_t *__restrict__ arena, uint32_t c0, uint32_t c1,
                                                    uint32_t file_c0, Sink sink) {
    static_assert(PF >= 2 && PF <= 6, "prefetch depth");
    const int lane = threadIdx.x & 31;
    const uint4 *base = reinterpret_cast<const uint4 *>(arena) + lane;
    bool carry_hdr = (fasta_backscan(arena, c0, file_c0, lane) & 1u) != 0;
    // slot[u] holds chunk (group base + u); after chunk cc is consumed its slot is refilled with chunk cc+PF.
    // Loads are clamped to chunk c1+1, which the arena's two NUL tail chunks keep in bounds.
    const uint32_t cmax = c1 + 1;
    uint4 slot[PF];
#pragma unroll
    for (int i = 0; i < PF; i++) slot[i] = __ldg(base + (size_t)min(c0 + i, cmax) * 32);
    Lane cur = decode16(slot[0]);
    auto emit = [&](uint32_t x) { sink(x << 2); };
    for (uint32_t cg = c0; cg < c1; cg += PF) {
#pragma unroll
        for (int u = 0; u < PF; u++) {
            const uint32_t c = cg + u;
            if (c < c1) {
                const uint4 wcur = slot[u];
                const Lane nxt = decode16(slot[(u + 1) % PF]);
                const unsigned any_dirty = __ballot_sync(FULL, cur.dirty);
                uint32_t st = 0;   // in_hdr | in_ls << 1 | next chunk starts inside a header << 2
                if (any_dirty | (unsigned)carry_hdr) {
                    const bool cls = (c == file_c0) || arena[(uint64_t)c * CHUNK - 1] == 0x0Au;
                    st = fasta_resolve(wcur, carry_hdr, cls, lane);
                }
                const uint32_t slowflag = cur.dirty | (st & 1u);
                const uint32_t xw = (cur.bits & ~3u) | slowflag;
                const uint32_t nx0 = (nxt.bits & ~3u) | nxt.dirty | ((st >> 2) & 1u);
                const uint32_t nbw = __shfl_sync(FULL, lane == 0 ? nx0 : xw, (lane + 1) & 31);
                const bool slow = FORCE_WALKER || slowflag || (nbw & 1u);
                if (!slow) {
                    const uint32_t nb = nbw & ~3u;
                    uint32_t hi = cur.bits, lo = nb;
                    if (cur.n == 15) { hi |= nb >> 30; lo = nb << 2; }
#pragma unroll
                    for (int j = 0; j < 15; j++) sink(kmer_off_at<K>(hi, lo, j));
                    if (cur.n == 16) sink(kmer_off_at<K>(hi, lo, 15));
                } else {
                    fasta_walk_lane<K>(arena, (uint64_t)c * CHUNK + (uint64_t)lane * 16, (st & 1u) != 0, (st & 2u) != 0, emit);
                }
                carry_hdr = (st & 4u) != 0;
                cur = nxt;
                slot[u] = __ldg(base + (size_t)min(c + PF, cmax) * 32);
            }
        }
    }
}

// Sinks: where a counted forward k-mer goes.
struct SmemSink {   // per-CTA privatised u32 histogram in shared memory
#ifdef KF_EMU
    uint32_t *hist;
    __device__ __forceinline__ void operator()(uint32_t off) const { atomicAdd(hist + (off >> 2), 1u); }
#else
    uint32_t base;  // shared-window byte address of bin 0
    __device__ __forceinline__ void operator()(uint32_t off) const {
        asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(base + off) : "memory");
    }
#endif
};
__device__ __forceinline__ SmemSink make_smem_sink(uint32_t *hist) {
    SmemSink s;
#ifdef KF_EMU
    s.hist = hist;
#else
    s.base = (uint32_t)__cvta_generic_to_shared(hist);
#endif
    return s;
}
struct GmemSink {   // dense u32 forward counts of one file in global memory (k >= 8)
    uint32_t *g;
    __device__ __forceinline__ void operator()(uint32_t off) const {
        atomicAdd(reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(g) + off), 1u);
    }
};

// ------------------------------------------------------------------------------------------------
// Counting kernels
// ------------------------------------------------------------------------------------------------
// k <= 7: per-CTA privatised u32 histogram in shared memory, flushed to the per-file u64 forward
// counts when the CTA moves to another file.  Persistent: CTA b owns tiles [cta_begin[b], cta_begin[b+1]).
template <int K, int THREADS, int MIN_CTAS, bool FORCE_WALKER, int PF = 3>
__global__ void __launch_bounds__(THREADS, MIN_CTAS)
count_fasta_smem_kernel(const uint8_t *__restrict__ arena, const Tile *__restrict__ tiles,
                        const int *__restrict__ cta_begin, unsigned long long *__restrict__ g_fwd) {
    KF_DYN_SMEM(uint32_t, hist);
    constexpr int NB = 1 << (2 * K);
    constexpr int NWARPS = THREADS / 32;
    for (int i = threadIdx.x; i < NB; i += THREADS) hist[i] = 0;
    __syncthreads();
    const int warp = threadIdx.x >> 5;
    int cur_file = -1;
    auto flush = [&](int file) {
        __syncthreads();
        unsigned long long *g = g_fwd + (size_t)file * NB;
        for (int i = threadIdx.x; i < NB; i += THREADS) {
            const uint32_t v = hist[i];
            if (v) { atomicAdd(g + i, (unsigned long long)v); hist[i] = 0; }
        }
        __syncthreads();
    };
    const SmemSink emit = make_smem_sink(hist);
    const int t1 = cta_begin[blockIdx.x + 1];
    for (int t = cta_begin[blockIdx.x]; t < t1; ++t) {
        const Tile T = tiles[t];
        if ((int)T.file != cur_file) {
            if (cur_file >= 0) flush(cur_file);
            cur_file = (int)T.file;
        }
        const uint32_t cpw = (T.n_chunks + NWARPS - 1) / NWARPS;
        const uint32_t c0 = T.first_chunk + (uint32_t)warp * cpw;
        const uint32_t cend = T.first_chunk + T.n_chunks;
        const uint32_t c1 = (c0 + cpw < cend) ? c0 + cpw : cend;
        if (c0 < c1) fasta_process_range<K, FORCE_WALKER, PF>(arena, c0, c1, T.file_chunk0, emit);
    }
    if (cur_file >= 0) flush(cur_file);
}

// k >= 8: forward counts live in global memory (u32 [file][4^k], L2-resident for k <= 10);
// every occurrence is one RED.  Same parser, different sink.
template <int K, int THREADS, bool FORCE_WALKER>
__global__ void __launch_bounds__(THREADS)
count_fasta_gmem_kernel(const uint8_t *__restrict__ arena, const Tile *__restrict__ tiles,
                        const int *__restrict__ cta_begin, uint32_t *__restrict__ g_fwd32, uint32_t file_base) {
    constexpr size_t NB = (size_t)1 << (2 * K);
    constexpr int NWARPS = THREADS / 32;
    const int warp = threadIdx.x >> 5;
    const int t1 = cta_begin[blockIdx.x + 1];
    for (int t = cta_begin[blockIdx.x]; t < t1; ++t) {
        const Tile T = tiles[t];
        GmemSink emit;
        emit.g = g_fwd32 + (size_t)(T.file - file_base) * NB;
        const uint32_t cpw = (T.n_chunks + NWARPS - 1) / NWARPS;
        const uint32_t c0 = T.first_chunk + (uint32_t)warp * cpw;
        const uint32_t cend = T.first_chunk + T.n_chunks;
        const uint32_t c1 = (c0 + cpw < cend) ? c0 + cpw : cend;
        if (c0 < c1) fasta_process_range<K, FORCE_WALKER, 3>(arena, c0, c1, T.file_chunk0, emit);
    }
}

// ------------------------------------------------------------------------------------------------
// Fold to canonical + total + pseudocount + normalise (main.py:327-342), one CTA per file
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t revcomp_std(uint32_t x, int k) {
    uint32_t y = __brev(~x);
    y = ((y & 0xAAAAAAAAu) >> 1) | ((y & 0x55555555u) << 1);
    return y >> (32 - 2 * k);
}
__device__ __forceinline__ uint32_t std_to_gray(uint32_t x) { return x ^ ((x >> 1) & 0x55555555u); }

template <typename FwdT>
__global__ void __launch_bounds__(1024)
fold_normalize_kernel(const FwdT *__restrict__ g_fwd, const uint32_t *__restrict__ canon, int k, long long V,
                      uint32_t flags, uint32_t file_base, unsigned long long *__restrict__ counts,
                      double *__restrict__ freq, float *__restrict__ feat,
                      unsigned long long *__restrict__ totals) {
    const size_t NB = (size_t)1 << (2 * k);
    const uint32_t file = blockIdx.x;
    const FwdT *g = g_fwd + (size_t)file * NB;
    const size_t orow = (size_t)(file + file_base) * (size_t)V;
    __shared__ unsigned long long red[32];
    __shared__ unsigned long long s_total;
    auto canon_count = [&](long long i) -> unsigned long long {
        const uint32_t m = canon[i];
        const uint32_t r = revcomp_std(m, k);
        unsigned long long c = (unsigned long long)g[std_to_gray(m)];
        if (r != m) c += (unsigned long long)g[std_to_gray(r)];
        return c;
    };
    unsigned long long local = 0;
    for (long long i = threadIdx.x; i < V; i += blockDim.x) {
        const unsigned long long c = canon_count(i);
        if (counts) counts[orow + i] = c;
        local += c;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(FULL, local, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x < 32) {
        unsigned long long v = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
        if (threadIdx.x == 0) { s_total = v; if (totals) totals[file + file_base] = v; }
    }
    __syncthreads();
    if (!freq && !feat) return;
    const bool pc = flags & 1u, raw = flags & 2u;
    const double denom = (double)s_total + (pc ? 0.5 * (double)V : 0.0);
    for (long long i = threadIdx.x; i < V; i += blockDim.x) {
        double v = (double)canon_count(i) + (pc ? 0.5 : 0.0);
        if (!raw) v = v / denom;   // IEEE fp64 division: correctly rounded, bit-exact with numpy
        if (freq) freq[orow + i] = v;
        if (feat) feat[orow + i] = (float)(v * 1e4);   // train_classifier_model.py:149,323
    }
}

}  // namespace kf
