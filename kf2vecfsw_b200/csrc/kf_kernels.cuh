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

// Where sequence bytes are read from.  GlobalSrc: the arena in HBM.  WindowSrc: the same, but bytes that lie in
// the window [w0, w0+wlen) the line-grid kernel has staged in shared memory are served from there, so its
// off-grid handling does not pay global round trips under a saturated memory system.
struct GlobalSrc {
    const uint8_t *arena;
    __device__ __forceinline__ uint4 load16(uint32_t chunk, int lane) const {
        return __ldg(reinterpret_cast<const uint4 *>(arena) + (size_t)chunk * 32 + lane);
    }
    __device__ __forceinline__ uint32_t byte(uint64_t p) const { return arena[p]; }
};
struct WindowSrc {
    const uint8_t *arena;
    const uint8_t *win;
    uint64_t w0;     // multiple of 16
    uint32_t wlen;
    __device__ __forceinline__ uint4 load16(uint32_t chunk, int lane) const {
        const uint64_t d = (uint64_t)chunk * CHUNK + (uint64_t)lane * 16 - w0;
        if (d <= (uint64_t)(wlen - 16)) return *reinterpret_cast<const uint4 *>(win + d);   // (d wraps when below w0)
        return __ldg(reinterpret_cast<const uint4 *>(arena) + (size_t)chunk * 32 + lane);
    }
    __device__ __forceinline__ uint32_t byte(uint64_t p) const {
        const uint64_t d = p - w0;
        return d < (uint64_t)wlen ? win[d] : arena[p];
    }
};

// Byte walker for one lane: counts every k-mer whose first base lies in [p0, p1); (in_hdr, at_ls) is the
// line state at p0.  Walks past p1 only as far as the k-mers it owns reach.
template <int K, class Src, class Emit>
__device__ KF_NOINLINE void fasta_walk_lane(const Src src, uint64_t p0, uint64_t p1, bool in_hdr, bool at_ls,
                                            Emit emit) {
    constexpr uint32_t MASK = (K >= 16) ? 0xFFFFFFFFu : ((1u << (2 * K)) - 1u);
    uint32_t kmer = 0;
    int run = 0, owned = 0;
    uint64_t p = p0;
    const uint64_t own_end = p1;
    for (;;) {
        const bool own = p < own_end;
        if (!own && (owned == 0 || run - K + 1 >= owned)) break;
        const uint32_t c = src.byte(p);
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
// Optional ownership clip [own_lo, own_hi) in arena bytes (own_lo must be a line start): only k-mers whose
// first base lies inside it are counted; lanes cut by a bound use the walker on their part.
// A sink may take a lane's whole 16-byte piece at once (window(hi, lo, n): the k-mers that start at bases 0..n-1 of the
// 32-base window hi:lo) instead of one k-mer at a time.
template <class S> struct sink_takes_window { static constexpr bool value = false; };
// What the rare paths (byte walker, out of line) count into: a by-value functor, so that the address of the sink proper
// is never taken and its state stays in registers on the hot path.  Default: a copy of the sink.
template <class S> __device__ __forceinline__ S sink_slow(const S &s) { return s; }
// A sink may want to see EVERY lane of every chunk (record(chunk * 32 + lane, fast, hi, lo, n): fast = the lane's k-mers
// go through window()/operator() of the fast path, otherwise through the rare paths or nowhere).
template <class S> struct sink_records_lanes { static constexpr bool value = false; };
// A sink may bring its own byte walker for the rare paths (walk(src, p0, p1, in_hdr, at_ls): same contract as
// fasta_walk_lane) -- the sparse path's k-mers are canonical 64-bit codes with a run-time k.
template <class S> struct sink_walks_itself { static constexpr bool value = false; };
template <int K, bool FORCE_WALKER, int PF, class Sink, class Src>
__device__ __forceinline__ void fasta_process_range(const Src src, uint32_t c0, uint32_t c1, uint32_t file_c0, Sink &sink,
                                                    uint64_t own_lo = 0, uint64_t own_hi = ~0ull,
                                                    bool skip_backscan = false) {
    static_assert(PF >= 2 && PF <= 6, "prefetch depth");
    const int lane = threadIdx.x & 31;
    // With an ownership clip that starts at a line start inside (or at the start of) chunk c0, the state at the
    // chunk's first byte does not matter: the '\n' before own_lo decides every owned lane's state.
    bool carry_hdr = skip_backscan ? false : (fasta_backscan(src.arena, c0, file_c0, lane) & 1u) != 0;
    // slot[u] holds chunk (group base + u); after chunk cc is consumed its slot is refilled with chunk cc+PF.
    // Loads are clamped to chunk c1+1, which the arena's two NUL tail chunks keep in bounds.
    const uint32_t cmax = c1 + 1;
    uint4 slot[PF];
#pragma unroll
    for (int i = 0; i < PF; i++) slot[i] = src.load16(min(c0 + i, cmax), lane);
    Lane cur = decode16(slot[0]);
    auto slow_sink = sink_slow(sink);
    auto emit = [slow_sink](uint32_t x) mutable { slow_sink(x << 2); };
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
                    const bool cls = (c == file_c0) || src.byte((uint64_t)c * CHUNK - 1) == 0x0Au;
                    st = fasta_resolve(wcur, carry_hdr, cls, lane);
                }
                const uint32_t slowflag = cur.dirty | (st & 1u);
                const uint32_t xw = (cur.bits & ~3u) | slowflag;
                const uint32_t nx0 = (nxt.bits & ~3u) | nxt.dirty | ((st >> 2) & 1u);
                const uint32_t nbw = __shfl_sync(FULL, lane == 0 ? nx0 : xw, (lane + 1) & 31);
                const uint64_t pb = (uint64_t)c * CHUNK + (uint64_t)lane * 16;
                const bool outside = pb + 16 <= own_lo || pb >= own_hi;
                const bool cut = pb < own_lo || pb + 16 > own_hi;
                const bool slow = FORCE_WALKER || slowflag || (nbw & 1u) || cut;
                const uint32_t nb = nbw & ~3u;
                uint32_t hi = cur.bits, lo = nb;
                if (cur.n == 15) { hi |= nb >> 30; lo = nb << 2; }
                if constexpr (sink_records_lanes<Sink>::value) sink.record((size_t)c * 32 + (size_t)lane, !outside && !slow, hi, lo, cur.n);
                if (outside) {
                } else if (!slow) {
                    if constexpr (sink_takes_window<Sink>::value) {
                        sink.window(hi, lo, cur.n);
                    } else {
#pragma unroll
                        for (int j = 0; j < 15; j++) sink(kmer_off_at<K>(hi, lo, j));
                        if (cur.n == 16) sink(kmer_off_at<K>(hi, lo, 15));
                    }
                } else if (pb < own_lo) {
                    if constexpr (sink_walks_itself<Sink>::value) sink.walk(src, own_lo, pb + 16 < own_hi ? pb + 16 : own_hi, false, true);
                    else fasta_walk_lane<K>(src, own_lo, pb + 16 < own_hi ? pb + 16 : own_hi, false, true, emit);
                } else {
                    if constexpr (sink_walks_itself<Sink>::value) sink.walk(src, pb, pb + 16 < own_hi ? pb + 16 : own_hi, (st & 1u) != 0, (st & 2u) != 0);
                    else fasta_walk_lane<K>(src, pb, pb + 16 < own_hi ? pb + 16 : own_hi, (st & 1u) != 0, (st & 2u) != 0, emit);
                }
                carry_hdr = (st & 4u) != 0;
                cur = nxt;
                slot[u] = src.load16(min(c + PF, cmax), lane);
            }
        }
    }
}

// Sinks: where a counted forward k-mer goes.
struct SmemSink {   // per-CTA privatised u32 histogram in shared memory
#ifdef KF_EMU
    uint32_t *hist;
    __device__ __forceinline__ void operator()(uint32_t off) const { atomicAdd(hist + (off >> 2), 1u); }
    __device__ __forceinline__ void add_if(uint32_t off, uint32_t t, uint32_t bit) const { if (t & bit) atomicAdd(hist + (off >> 2), 1u); }
#else
    uint32_t base;  // shared-window byte address of bin 0
    __device__ __forceinline__ void operator()(uint32_t off) const {
        asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(base + off) : "memory");
    }
    // count iff (t & bit) != 0: one predicated RED, no branch
    __device__ __forceinline__ void add_if(uint32_t off, uint32_t t, uint32_t bit) const {
        asm volatile("{\n\t.reg .pred p;\n\t.reg .b32 a;\n\tand.b32 a, %1, %2;\n\tsetp.ne.u32 p, a, 0;\n\t@p red.shared.add.u32 [%0], 1;\n\t}" ::"r"(base + off),
                     "r"(t), "r"(bit)
                     : "memory");
    }
#endif
};
template <class S> struct sink_has_add_if { static constexpr bool value = false; };
template <> struct sink_has_add_if<SmemSink> { static constexpr bool value = true; };
template <> struct sink_has_add_if<const SmemSink> { static constexpr bool value = true; };
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
                        const int *__restrict__ cta_begin, unsigned long long *__restrict__ g_fwd,
                        const uint32_t *__restrict__ file_row, const uint32_t *__restrict__ file_P,
                        const uint32_t *__restrict__ width_counts) {
    KF_DYN_SMEM(uint32_t, hist);
    constexpr int NB = 1 << (2 * K);
    constexpr int NWARPS = THREADS / 32;
    if (width_counts && width_counts[0] == 0) return;   // every file went to a line-grid launch
    for (int i = threadIdx.x; i < NB; i += THREADS) hist[i] = 0;
    __syncthreads();
    const int warp = threadIdx.x >> 5;
    int cur_file = -1;
    auto flush = [&](int file) {
        __syncthreads();
        unsigned long long *g = g_fwd + (size_t)file_row[file] * NB;   // the file's first row (rows are summed by the fold)
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
        if (file_P && file_P[T.file] != 0) continue;   // taken by a line-grid launch
        if ((int)T.file != cur_file) {
            if (cur_file >= 0) flush(cur_file);
            cur_file = (int)T.file;
        }
        const uint32_t cpw = (T.n_chunks + NWARPS - 1) / NWARPS;
        const uint32_t c0 = T.first_chunk + (uint32_t)warp * cpw;
        const uint32_t cend = T.first_chunk + T.n_chunks;
        const uint32_t c1 = (c0 + cpw < cend) ? c0 + cpw : cend;
        if (c0 < c1) fasta_process_range<K, FORCE_WALKER, PF>(GlobalSrc{arena}, c0, c1, T.file_chunk0, emit);
    }
    if (cur_file >= 0) flush(cur_file);
}

// k >= 8: forward counts live in global memory (u32 [file][4^k], L2-resident for k <= 10);
// every occurrence is one RED.  Same parser, different sink.
template <int K, int THREADS, bool FORCE_WALKER>
__global__ void __launch_bounds__(THREADS)
count_fasta_gmem_kernel(const uint8_t *__restrict__ arena, const Tile *__restrict__ tiles,
                        const int *__restrict__ cta_begin, uint32_t *__restrict__ g_fwd32, uint32_t file_base,
                        const uint32_t *__restrict__ file_skip) {
    constexpr size_t NB = (size_t)1 << (2 * K);
    constexpr int NWARPS = THREADS / 32;
    const int warp = threadIdx.x >> 5;
    const int t1 = cta_begin[blockIdx.x + 1];
    for (int t = cta_begin[blockIdx.x]; t < t1; ++t) {
        const Tile T = tiles[t];
        if (file_skip && file_skip[T.file]) continue;   // counted by the partitioned kernel
        GmemSink emit;
        emit.g = g_fwd32 + (size_t)(T.file - file_base) * NB;
        const uint32_t cpw = (T.n_chunks + NWARPS - 1) / NWARPS;
        const uint32_t c0 = T.first_chunk + (uint32_t)warp * cpw;
        const uint32_t cend = T.first_chunk + T.n_chunks;
        const uint32_t c1 = (c0 + cpw < cend) ? c0 + cpw : cend;
        if (c0 < c1) fasta_process_range<K, FORCE_WALKER, 3>(GlobalSrc{arena}, c0, c1, T.file_chunk0, emit);
    }
}

// ================================================================================================
// Line kernel (k = 7): fixed-width FASTA, one line per lane -- shared pieces
// ================================================================================================
// Most FASTA is written with a fixed line width (NCBI 80, Ensembl 60, UCSC/others 70).  When a warp knows
// the width LW it gives every lane one whole line: the '\n' sits at a compile-time byte, so nothing has to
// be searched or deleted, lanes never exchange data (each lane also decodes the first 6 bases of the next
// line as look-ahead), and LW being even lets a lane count its LW 7-mers as LW/2 8-mers ("pairs": the
// 8-mer at base 2i is the 7-mer at 2i followed by the one at 2i+1).  Halving the shared atomics matters
// because their bank-conflict wavefronts are the first limiter of the generic kernel (profiles/r01_*).
//
//   staging  : per-warp buffer in shared memory filled by one TMA bulk copy (cp.async.bulk + mbarrier)
//              of 32 lines.
//   histogram: 65,536 8-mer bins as 32,768 u32 words: word = v >> 1; the low half counts every pair of
//              the word, the high half those with v & 1 (addend 1 or 0x10001).  A half can wrap, so at
//              every flush the sum of the low halves is compared with the number of pairs issued; on a
//              mismatch the CTA discards the histogram and recounts its part of that file with the
//              generic path into global memory (exact, slow, practically never taken).
//   irregular: lines that break the grid (record ends, headers, other widths) and lanes with non-ACGT
//              bytes go through the generic range processor / byte walker with a global-memory sink.

#ifdef KF_EMU
#define KF_PREFETCH_L2(p) ((void)(p))
__device__ inline void __syncwarp_emu() { emu::exchange(0, 0); }
#define KF_SYNCWARP() __syncwarp_emu()
#define KF_LDCG(p) (*(p))
__device__ inline void stage_bar_init(uint64_t *) {}
__device__ inline void stage_issue(void *dst, const void *src, uint32_t bytes, uint64_t *) { memcpy(dst, src, bytes); }
__device__ inline void stage_issue2(void *d0, const void *s0, void *d1, const void *s1, uint32_t bytes, uint64_t *) { memcpy(d0, s0, bytes); memcpy(d1, s1, bytes); }
__device__ inline void stage_wait(uint64_t *, uint32_t) { KF_SYNCWARP(); }
__device__ inline uint32_t smem_addr(const void *) { return 0; }
__device__ inline void red_shared_add(uint32_t *hist, uint32_t, uint32_t off, uint32_t v) { atomicAdd(hist + (off >> 2), v); }
__device__ inline void red_shared_add_if(uint32_t *hist, uint32_t, uint32_t off, uint32_t v, uint32_t t, uint32_t bit) { if (t & bit) atomicAdd(hist + (off >> 2), v); }
template <uint32_t BASE> __device__ inline void red_shared_add_imm(uint32_t *hist, uint32_t, uint32_t off, uint32_t v) { atomicAdd(hist + (off >> 2), v); }
__device__ inline uint32_t opaque_one() { return gridDim.y; }
#else
#define KF_PREFETCH_L2(p) asm volatile("prefetch.global.L2 [%0];" ::"l"(p))
#define KF_SYNCWARP() __syncwarp()
#define KF_LDCG(p) __ldcg(p)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void stage_bar_init(uint64_t *bar) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// one lane: arm the barrier with the byte count and start the bulk copy global -> shared (TMA)
__device__ __forceinline__ void stage_issue(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic reads of dst vs async write
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// one lane: two bulk copies of `bytes` each on one barrier phase
__device__ __forceinline__ void stage_issue2(void *d0, const void *s0, void *d1, const void *s1, uint32_t bytes, uint64_t *bar) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(2u * bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(d0)),
                 "l"(s0), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(d1)),
                 "l"(s1), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void stage_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "KF_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra KF_DONE;\n\t"
        "bra KF_WAIT;\n\t"
        "KF_DONE:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return smem_u32(p); }
// RED iff (t & bit) != 0, as a predicated instruction (no branch, no reconvergence barrier around it)
__device__ __forceinline__ void red_shared_add_if(uint32_t *, uint32_t base, uint32_t off, uint32_t v, uint32_t t, uint32_t bit) {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b32 a;\n\tand.b32 a, %2, %3;\n\tsetp.ne.u32 p, a, 0;\n\t@p red.shared.add.u32 [%0], %1;\n\t}" ::"r"(base + off),
                 "r"(v), "r"(t), "r"(bit)
                 : "memory");
}
__device__ __forceinline__ void red_shared_add(uint32_t *, uint32_t base, uint32_t off, uint32_t v) {
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(base + off), "r"(v) : "memory");
}
// The same with the histogram's shared-window address as an IMMEDIATE of the instruction (BASE != 0), saving the add per
// RED; BASE == 0: address = base + off as above.  The caller has checked that the histogram really sits at BASE.
template <uint32_t BASE>
__device__ __forceinline__ void red_shared_add_imm(uint32_t *, uint32_t base, uint32_t off, uint32_t v) {
    if (BASE != 0) asm volatile("red.shared.add.u32 [%0+%2], %1;" ::"r"(off), "r"(v), "n"(BASE) : "memory");
    else asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(base + off), "r"(v) : "memory");
}
// the constant 1 in a register the compiler cannot see through (every launch of this library is one-dimensional, so
// gridDim.y == 1): (x & IMM) | one is then ONE LOP3 -- an instruction takes one immediate; with two literal constants
// the compiler emits two
__device__ __forceinline__ uint32_t opaque_one() { return gridDim.y; }
#endif

// ------------------------------------------------------------------------------------------------
// k = 8..10: partitioned shared-memory histogram
// ------------------------------------------------------------------------------------------------
// 4^k counters do not fit in shared memory, and one RED per occurrence into an L2-resident row runs at the L2's atomic
// rate (~0.1 Tbases/s).  Instead the k-mer space is cut into partitions that fit in shared memory as u16 halves (PartGeom:
// 65,536 bins at k = 8, 98,304 at k = 9 and 10); a work item is (file, partition): the CTA counts the k-mers of its
// partition over the whole file and owns that part of the file's row outright -- plain read-add-write, no global atomics.
//   word w of the histogram: low half = occurrences of bins 2w and 2w+1, high half = those of bin 2w+1 (one RED with
//   addend 1 or 0x10001).  A half can wrap: the histogram is drained every PART_FLUSH_TILES tiles, and at every drain
//   the sum of the low halves must equal the number of REDs issued; if not, the interval is recounted with global REDs.
constexpr int PART_FLUSH_TILES = 8;

// Partition geometry.  TB = 0: one partition of 4^K bins (k = 8).  TB > 0 (odd): the k-mer's top TB bits -- its first
// TB/2 bases and the high code bit of the next -- number 2^TB "granules" of 4^K >> TB bins; a partition is three
// consecutive granules (98,304 bins = 192 KB of u16 halves at k = 9, 10: three passes instead of four at k = 9, eleven
// instead of sixteen at k = 10), the last one what is left.
template <int K, int TB>
struct PartGeom {
    static_assert(TB == 0 || (TB % 2 == 1 && TB >= 3 && TB <= 5), "top bits");
    static constexpr uint32_t GSH = 2 * K - TB;
    static constexpr uint32_t GRAN = 1u << GSH;
    static constexpr uint32_t NTOP = 1u << TB;
    static constexpr uint32_t NBINS = TB ? 3u * GRAN : GRAN;          // bins of a full partition
    static constexpr uint32_t NPART = TB ? (NTOP + 2u) / 3u : 1u;
    static constexpr uint32_t NWORDS = NBINS / 2;
    static constexpr uint32_t AMASK = (TB ? 8u * GRAN : 2u * GRAN) - 4u;   // byte address of the word from 2 * (kmer - first bin)
    static_assert(NBINS >= 8 && NWORDS * 4 <= 196608, "partition size");
    static __host__ __device__ constexpr uint32_t first_bin(uint32_t p) { return TB ? 3u * p * GRAN : 0u; }
    static __host__ __device__ constexpr uint32_t bins_of(uint32_t p) {
        return TB ? ((1u << (2 * K)) - first_bin(p) < NBINS ? (1u << (2 * K)) - first_bin(p) : NBINS) : NBINS;
    }
    static __device__ __forceinline__ bool owns(uint32_t kmer, uint32_t p) { return TB == 0 || ((kmer >> GSH) - 3u * p) < 3u; }
};

template <int K, int TB>
struct PartSink {
    using G = PartGeom<K, TB>;
    static constexpr uint32_t NWORDS = G::NWORDS;
    uint32_t *hist;
    uint32_t base;     // shared-window address of word 0
    uint32_t part;     // partition index
    uint32_t lo2;      // 2 * first bin of the partition
    uint32_t issued;   // REDs issued by this thread since the last drain
    __device__ __forceinline__ void set_part(uint32_t p) { part = p; lo2 = 2u * G::first_bin(p); }
    // sv: the k-mer on bits 2K:1 (anything above)
    __device__ __forceinline__ uint32_t addr_of(uint32_t sv) const { return (TB ? sv - lo2 : sv) & G::AMASK; }
    __device__ __forceinline__ void add(uint32_t sv) { red_shared_add(hist, base, addr_of(sv), (sv & 2u) * 0x8000u + 1u); }
    uint32_t *s_slow;  // shared: REDs issued by the rare paths since the last drain
    uint32_t *g_row;   // pass A of a multi-pass count: the file's row, for the rare paths' k-mers of OTHER partitions
    uint2 *stream;     // pass A: every lane's decoded piece goes here for the passes over the other partitions
    // stream entry: x = hi, y = lo with bit 0 = "not a fast lane" (nothing to count from the stream), bit 1 = 15 k-mers
    // (lo's low two bits are never part of a k-mer: the last one, j = 15, ends at base 24 for K = 10)
    __device__ __forceinline__ void record(size_t idx, bool fast, uint32_t hi, uint32_t lo, uint32_t n) const {
        if (stream) stream[idx] = make_uint2(hi, (lo & ~3u) | (fast ? 0u : 1u) | (n == 15 ? 2u : 0u));
    }
    // 0x55555555-style masks over the 16 positions of a window: bit 30 - 2j belongs to position j
    static __device__ __forceinline__ uint32_t match(uint32_t w, uint32_t c) {       // base j of w equals code c
        const uint32_t x = w ^ (c * 0x55555555u);
        return ~(x | (x >> 1)) & 0x55555555u;
    }
    // positions whose k-mer starts in one of this partition's granules
    __device__ __forceinline__ uint32_t own_mask(uint32_t hi, uint32_t lo, uint32_t n) const {
        uint32_t t = 0x55555555u;
        if (TB) {
            constexpr int NBASE = TB / 2;
            uint32_t w[NBASE + 1];
#pragma unroll
            for (int b = 0; b <= NBASE; b++) w[b] = b == 0 ? hi : __funnelshift_l(lo, hi, 2 * b);
            const uint32_t hb = (w[NBASE] >> 1) & 0x55555555u;   // high code bit of base j + NBASE
            t = 0;
#pragma unroll
            for (uint32_t q = 0; q < 3; q++) {
                const uint32_t T = 3u * part + q;   // granule number: NBASE bases and one bit
                uint32_t m = (T & 1u) ? hb : (~hb & 0x55555555u);
#pragma unroll
                for (int b = 0; b < NBASE; b++) m &= match(w[b], (T >> (TB - 2 - 2 * b)) & 3u);
                if (T < G::NTOP) t |= m;
            }
        }
        if (n == 15) t &= ~1u;
        if (n == 0) t = 0u;
        return t;
    }
    __device__ __forceinline__ void window(uint32_t hi, uint32_t lo, uint32_t n) {
        const uint32_t t = own_mask(hi, lo, n);
        issued += (uint32_t)__popc(t);
        // (walking the set bits of t instead was measured slower in the text pass and in the stream passes at four
        //  partitions: the divergent loop costs more than 16 predicated REDs)
#pragma unroll
        for (int j = 0; j < 16; j++) {
            const int r = 63 - 2 * K - 2 * j;   // k-mer j on bits 2K:1 of (hi:lo) >> r
            const uint32_t sv = r >= 32 ? (hi >> (r - 32)) : __funnelshift_r(lo, hi, r);
            if (TB == 0 && j < 15) add(sv);
            else red_shared_add_if(hist, base, addr_of(sv), (sv & 2u) * 0x8000u + 1u, t, 1u << (30 - 2 * j));
        }
    }
    // The same for sparse matches (eleven partitions: one k-mer in ten is ours): SLOTS straight-line "next set bit"
    // slots, each one predicated RED, instead of 16 predicated positions; a lane with more matches finishes in a loop
    // that the warp enters only then.  Every lane of the warp must call it (ballot); n = 0: nothing to count.
    template <int SLOTS>
    __device__ __forceinline__ void window_slots(uint32_t hi, uint32_t lo, uint32_t n) {
        uint32_t t = own_mask(hi, lo, n);
        issued += (uint32_t)__popc(t);
        auto take = [&](uint32_t have) {
            const uint32_t b = 31u - (uint32_t)__clz((int)(t | 1u));   // highest set bit = 30 - 2j (0 when t is empty)
            t &= ~(1u << b);
            const uint32_t r = 33u - 2u * K + b;                          // 63 - 2K - 2j
            const uint32_t sv = r >= 32u ? (hi >> (r - 32u)) : __funnelshift_r(lo, hi, r & 31u);
            red_shared_add_if(hist, base, addr_of(sv), (sv & 2u) * 0x8000u + 1u, have, 1u);
        };
#pragma unroll
        for (int sidx = 0; sidx < SLOTS; sidx++) take(t != 0u ? 1u : 0u);
        if (__ballot_sync(FULL, t != 0u) != 0u) {
            while (t) take(1u);
        }
    }
};
template <int K, int TB> struct sink_takes_window<PartSink<K, TB>> { static constexpr bool value = true; };
template <int K, int TB> struct sink_records_lanes<PartSink<K, TB>> { static constexpr bool value = TB != 0; };
template <int K, int TB>
struct PartSlowSink {   // one k-mer at a time (rare paths): off = 4 * kmer
    uint32_t *hist;
    uint32_t base, part, lo2;
    uint32_t *s_slow;
    uint32_t *g_row;
    __device__ __forceinline__ void operator()(uint32_t off) const {
        if (PartGeom<K, TB>::owns(off >> 2, part)) {
            const uint32_t sv = off >> 1;
            red_shared_add(hist, base, (TB ? sv - lo2 : sv) & PartGeom<K, TB>::AMASK, (sv & 2u) * 0x8000u + 1u);
            atomicAdd(s_slow, 1u);
        } else if (g_row) {
            atomicAdd(reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(g_row) + off), 1u);   // another partition's bin
        }
    }
};
template <int K, int TB> __device__ __forceinline__ PartSlowSink<K, TB> sink_slow(const PartSink<K, TB> &s) {
    return PartSlowSink<K, TB>{s.hist, s.base, s.part, s.lo2, s.s_slow, s.g_row};
}
// exact recount of a stream entry's k-mers of one partition: global REDs
template <int K, int TB>
__device__ __forceinline__ void part_stream_recount(uint32_t hi, uint32_t lo, uint32_t n, uint32_t part, uint32_t *g_row) {
    const unsigned long long w = ((unsigned long long)hi << 32) | lo;
    for (uint32_t j = 0; j < n; j++) {
        const uint32_t kmer = (uint32_t)(w >> (64 - 2 * K - 2 * j)) & ((1u << (2 * K)) - 1u);
        if (PartGeom<K, TB>::owns(kmer, part)) atomicAdd(g_row + kmer, 1u);
    }
}

template <int K, int TB>
struct PartGmemSink {   // exact recount of one partition after a wrapped half: one global RED per occurrence
    uint32_t *g;        // the file's row
    uint32_t part;
    __device__ __forceinline__ void operator()(uint32_t off) const {
        if (PartGeom<K, TB>::owns(off >> 2, part)) atomicAdd(reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(g) + off), 1u);
    }
};

// items: (file, partition) pairs, the partitions of a file next to each other; CTAs take them from a global counter.
// file_t0[f] .. file_t0[f + 1]: the file's tiles in `tiles` (consecutive); files with no tiles or with small == 1 are
// left to the global-atomic kernel.  Rows are zero when the kernel starts.
// MODE 0: the only pass (one partition, k = 8).  MODE 1: pass A of a multi-pass count -- the text is parsed ONCE: the
// item's partition is counted, every lane's decoded piece is written to `stream` (8 bytes per 16 bytes of text), and
// the k-mers of the rare paths that belong to other partitions go straight to the row with global REDs.  MODE 2: pass B
// -- the other partitions are counted from the stream (no parsing, half the bytes); launched after pass A.
template <int K, int TB, int THREADS, int MODE>
__global__ void __launch_bounds__(THREADS, 1)
count_fasta_part_kernel(const uint8_t *__restrict__ arena, const Tile *__restrict__ tiles, const int *__restrict__ file_t0,
                        const uint32_t *__restrict__ items, int n_items, uint32_t *__restrict__ g_fwd32, uint32_t file_base,
                        unsigned int *__restrict__ item_counter, uint2 *__restrict__ stream) {
    using S = PartSink<K, TB>;
    using PG = PartGeom<K, TB>;
    constexpr int NWARPS = THREADS / 32;
    constexpr size_t NB = (size_t)1 << (2 * K);
    KF_DYN_SMEM(uint32_t, hist);
    __shared__ unsigned long long s_red[2 * (THREADS / 32)];
    __shared__ int s_item;
    __shared__ uint32_t s_slow;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (uint32_t i = threadIdx.x; i < S::NWORDS; i += THREADS) hist[i] = 0;
    if (threadIdx.x == 0) s_slow = 0;
    S sink;
    sink.hist = hist;
    sink.base = smem_addr(hist);
    sink.issued = 0;
    sink.s_slow = &s_slow;
    sink.g_row = nullptr;
    sink.stream = MODE == 1 ? stream : nullptr;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_item = (int)atomicAdd(item_counter, 1u);
        __syncthreads();
        const int it = s_item;
        if (it >= n_items) break;
        const uint32_t file = items[it] >> 8, part = items[it] & 0xFFu;
        sink.set_part(part);
        uint32_t *file_row = g_fwd32 + (size_t)(file - file_base) * NB;
        uint32_t *row = file_row + PG::first_bin(part);
        const uint32_t nwords = PG::bins_of(part) / 2;   // (the last partition may be shorter)
        if (MODE == 1) sink.g_row = file_row;
        const int t0 = file_t0[file], t1 = file_t0[file + 1];
        for (int ta = t0; ta < t1; ta += PART_FLUSH_TILES) {
            const int tb = ta + PART_FLUSH_TILES < t1 ? ta + PART_FLUSH_TILES : t1;
            for (int t = ta; t < tb; ++t) {
                const Tile T = tiles[t];
                const uint32_t cpw = (T.n_chunks + NWARPS - 1) / NWARPS;
                const uint32_t c0 = T.first_chunk + (uint32_t)warp * cpw;
                const uint32_t cend = T.first_chunk + T.n_chunks;
                const uint32_t c1 = (c0 + cpw < cend) ? c0 + cpw : cend;
                if (MODE != 2) {
                    if (c0 < c1) fasta_process_range<K, false, 3>(GlobalSrc{arena}, c0, c1, T.file_chunk0, sink);
                } else {
                    // pass B: the lanes' decoded pieces, D chunks in flight (register ring, rotated at compile time)
                    constexpr int D = 4;
                    const uint2 *sp = stream + (size_t)c0 * 32 + lane;
                    uint2 ring[D];
#pragma unroll
                    for (int i = 0; i < D; i++) ring[i] = c0 + i < c1 ? KF_LDCG(sp + (size_t)i * 32) : make_uint2(0u, 1u);
                    for (uint32_t cg = c0; cg < c1; cg += D) {
#pragma unroll
                        for (int u = 0; u < D; u++) {
                            const uint32_t c = cg + u;
                            const uint2 v = ring[u];
                            ring[u] = c + D < c1 ? KF_LDCG(sp + (size_t)(c + D - c0) * 32) : make_uint2(0u, 1u);
                            // (past c1: flagged entries; every lane takes part in window_slots' ballot)
                            // (measured: 4 slots 19.0 ms vs 16 positions 23.4 ms at k = 10; 8 slots at k = 9 are slower than 16 positions)
                            if (TB >= 5) sink.template window_slots<5>(v.x, (v.y & 1u) ? 0u : (v.y & ~3u), (v.y & 1u) ? 0u : ((v.y & 2u) ? 15u : 16u));
                            else if (!(v.y & 1u)) sink.window(v.x, v.y & ~3u, (v.y & 2u) ? 15u : 16u);
                        }
                    }
                }
            }
            // ---- drain: checksum, then add the interval's counts to the row ----
            unsigned long long iss = sink.issued;
            sink.issued = 0;
            __syncthreads();
            unsigned long long low = 0;
            for (uint32_t i = threadIdx.x; i < S::NWORDS; i += THREADS) low += hist[i] & 0xFFFFu;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { iss += __shfl_xor_sync(FULL, iss, o); low += __shfl_xor_sync(FULL, low, o); }
            if (lane == 0) { s_red[warp] = iss; s_red[NWARPS + warp] = low; }
            __syncthreads();
            unsigned long long ti = s_slow, tl = 0;
#pragma unroll
            for (int w = 0; w < NWARPS; w++) { ti += s_red[w]; tl += s_red[NWARPS + w]; }
            const bool ok = ti == tl;
            __syncthreads();
            if (threadIdx.x == 0) s_slow = 0;
            for (uint32_t i = threadIdx.x; i < S::NWORDS; i += THREADS) {
                const uint32_t w = hist[i];
                hist[i] = 0;
                if (ok && w && i < nwords) {
                    uint2 *rp = reinterpret_cast<uint2 *>(row) + i;
                    uint2 v = KF_LDCG(rp);   // (not through L1: an exact recount adds to these bins with REDs)
                    v.x += (w & 0xFFFFu) - (w >> 16);
                    v.y += w >> 16;
                    *rp = v;
                }
            }
            if (!ok) {
                // a half wrapped: this interval again, straight into the row (other CTAs never touch this partition's bins,
                // but this CTA's threads now share them: REDs)
                __threadfence();
                __syncthreads();
                PartGmemSink<K, TB> gs;
                gs.g = g_fwd32 + (size_t)(file - file_base) * NB;
                gs.part = part;
                for (int t = ta; t < tb; ++t) {
                    const Tile T = tiles[t];
                    const uint32_t cpw = (T.n_chunks + NWARPS - 1) / NWARPS;
                    const uint32_t c0 = T.first_chunk + (uint32_t)warp * cpw;
                    const uint32_t cend = T.first_chunk + T.n_chunks;
                    const uint32_t c1 = (c0 + cpw < cend) ? c0 + cpw : cend;
                    if (MODE != 2) {
                        if (c0 < c1) fasta_process_range<K, false, 2>(GlobalSrc{arena}, c0, c1, T.file_chunk0, gs);
                    } else {
                        for (uint32_t c = c0; c < c1; ++c) {
                            const uint2 v = KF_LDCG(stream + (size_t)c * 32 + lane);
                            if (!(v.y & 1u)) part_stream_recount<K, TB>(v.x, v.y & ~3u, (v.y & 2u) ? 15u : 16u, part, gs.g);
                        }
                    }
                }
                __threadfence();
            }
            __syncthreads();   // the histogram is empty again: the next interval's REDs may start
        }
    }
}

// Sinks of the line kernel's rare paths (byte walker, off-grid lines).  Both take off = 4 * kmer with the first base
// most significant and store it in the pair histogram's orientation (7-mer digits reversed: first base in bits 1:0).
__device__ __forceinline__ uint32_t rev7_of_off(uint32_t off) {   // off = 4 * kmer, 14-bit kmer
    const uint32_t r = __brev(off) >> 16;                          // bit-reverse, then swap the two bits of each digit
    return ((r & 0x2AAAu) >> 1) | ((r & 0x1555u) << 1);
}
struct SingleSink {   // "singles" histogram in shared memory: 16,384 7-mer bins as 8,192 words of two u16 halves
                      // (word = xk >> 1, half = xk & 1) plus the number of k-mers issued, for the overflow checksum
    uint32_t *h;
    uint32_t *n;
    __device__ __forceinline__ void operator()(uint32_t off) const {
        const uint32_t idx = rev7_of_off(off);
        atomicAdd(h + (idx >> 1), (idx & 1u) ? 0x10000u : 1u);
        atomicAdd(n, 1u);
#ifdef KF_EMU_DEBUG
        extern std::atomic<long> g_dbg_emits;
        g_dbg_emits++;
#endif
    }
};
struct Gmem64Sink {   // exact recount after a wrapped 16-bit half: straight into the CTA's u64 row in global memory
    unsigned long long *g;
    __device__ __forceinline__ void operator()(uint32_t off) const { atomicAdd(g + rev7_of_off(off), 1ull); }
};

// First line start at or after byte q (a line start is file_begin or the byte after a '\n'); file_end if none.
template <class Src>
__device__ KF_NOINLINE uint64_t fasta_line_start_at_or_after(const Src src, uint64_t q, uint64_t file_begin,
                                                             uint64_t file_end, int lane) {
    if (q <= file_begin) return file_begin;
    if (q >= file_end) return file_end;
    const uint64_t from = q - 1;   // a '\n' at q-1 makes q itself a line start
    for (uint64_t c = from / CHUNK; c * CHUNK < file_end; ++c) {
        const uint4 w = src.load16((uint32_t)c, lane);
        const uint64_t pb = c * CHUNK + (uint64_t)lane * 16;
        int first = 16;
#pragma unroll
        for (int i = 15; i >= 0; i--)
            if (byte_of(w, i) == 0x0Au && pb + i >= from) first = i;
        const unsigned Bm = __ballot_sync(FULL, first < 16);
        if (Bm) {
            const int j = __ffs((int)Bm) - 1;
            const int fj = __shfl_sync(FULL, first, j);
            const uint64_t r = c * CHUNK + (uint64_t)j * 16 + (uint64_t)fj + 1;
            return r < file_end ? r : file_end;
        }
    }
    return file_end;
}

// Exact but slow: the generic range processor over the lines starting in [lo, hi), global sink.
template <int K, class Src, class Sink>
__device__ KF_NOINLINE void lg_generic_region(const Src src, uint64_t lo, uint64_t hi, uint32_t file_c0, Sink &gs) {
    if (lo >= hi) return;
    const uint32_t c0 = (uint32_t)(lo / CHUNK), c1 = (uint32_t)((hi + CHUNK - 1) / CHUNK);
    fasta_process_range<K, false, 2>(src, c0, c1, file_c0, gs, lo, hi, true);
}

// ================================================================================================
// The kernel.  Its first generation (static split of a file piece over the warps, double-buffered windows,
// rare paths served from the staged window) was rebuilt around what its profile showed
// (profiles/r01_*): 15 % of the warp samples sat in the per-file flush barrier (static split of a file piece
// over the warps + rare-path detours of tens of microseconds), 10 % waited on global loads of those detours,
// and the ALU pipe (LOP3/SHF/PRMT, 64 lanes/clk/SM) was the busiest unit while the FMA pipe idled.
//   scheduling : a file piece is handed out to the warps in byte ranges ("units") from a shared cursor, large
//                first and shrinking towards the end of the piece (guided self-scheduling); unit sizes are
//                multiples of a window, anchored at the piece's first line start, so while the file stays on
//                one grid phase every unit is a whole number of full windows.  The next unit is claimed, and the
//                TMA copy of its first window started, before the current window is decoded.
//   staging    : one buffer per warp; the lane's line is pulled into registers, then the buffer is re-armed for
//                the next window at once (the copy overlaps the decode/count of this one).
//   decode     : codes are packed LSB-first ((w & 0x06060606) * M: no shift), validity is checked with left
//                shifts (IMAD, FMA pipe) and two LOP3 per word; k-mer indices are digit-reversed, undone at flush.
//   rare paths : a line holding a non-ACGT byte (or whose 6 look-ahead bytes do) goes to the exact byte walker on
//                global memory; a line that breaks the grid (short last line, header, other width) goes to the
//                exact generic range processor up to the next sequence line, where the grid restarts.
constexpr int LN_MAX_LW = 100;     // widest wrapped FASTA with a line-kernel instantiation (shared memory: 120 columns would not leave room)
constexpr int LN_SCR_WORDS = 36;   // per-warp scratch: one line + look-ahead re-fetched from global memory (P + LA + 3 bytes at LW = 120: 130)
template <int LW>
struct LineGeom {
    static_assert(LW % 2 == 0 && LW >= 32 && LW <= 120, "line width");
    static constexpr int P = LW + 1;                 // bytes per line incl. '\n'
    static constexpr int LA = 6;                     // look-ahead bases (K - 1)
    static constexpr int NBASES = LW + LA;
    static constexpr int NWA = (P + LA + 3) / 4;     // aligned words a lane decodes
    static constexpr int NPK = (NBASES + 15) / 16;   // packed registers
    static constexpr int NPAIR = LW / 2;
    static constexpr int WIN = 32 * P;               // bytes per full window
    static constexpr int NEED = 31 * P + 4 * NWA + 4;                 // bytes a window must hold from its first line start
    // DUAL: a window is staged as two halves of 16 lines (two bulk copies; 16 P is a multiple of 16), the second one HOFF
    // bytes after the first with HOFF = 64 (mod 128), i.e. 16 banks out of phase.  At P = 71 the lanes' first words fall
    // on 8 banks (17.75 words per line: four-way conflicts on every LDS); with the upper 16 lanes moved by 16 banks they
    // are two-way: 0.39 -> 0.45 of the roofline on 70-column files.  At P = 81 (two-way to start with) the same layout plus
    // windows that start at a multiple of 4 bytes is conflict-free, but measured SLOWER (round 1 and again in round 2:
    // one-record file 11.8 -> 10.8 bases/clk/SM, headline kernel 1.76 -> 1.81 ms): two bulk copies per window cost more
    // than the 29 wavefronts per window they save.
    static constexpr bool DUAL = LW == 70;
    static constexpr int HNEED = 15 * P + 4 * NWA + 4;                // bytes a half must hold from its first line start
    static constexpr int HCOPY = ((HNEED + 15 + 112 + 15) / 16) * 16; // + alignment + slack to find a unit's first line start
    static constexpr int HOFF = ((HCOPY + 63) / 128) * 128 + 64;      // >= HCOPY, = 64 (mod 128)
    static constexpr int STAGE = DUAL ? HOFF + HCOPY : ((NEED + 16 + 112 + 15) / 16) * 16;  // + alignment + slack to find a unit's first line start
    static_assert(!DUAL || (HOFF >= HCOPY && HOFF % 128 == 64 && (16 * P) % 16 == 0), "half layout");
    // staged bytes from `base` that line l of the window may use (its slot and look-ahead must lie inside the copy)
    static constexpr int WOFF_MAX = DUAL ? HCOPY - HNEED : STAGE - NEED;
};

__device__ __forceinline__ uint32_t digit_rev7(uint32_t x) {   // reverse the seven 2-bit digits of a 14-bit value
    const uint32_t r = __brev(x) >> 18;
    return ((r & 0x2AAAu) >> 1) | ((r & 0x1555u) << 1);
}

// Decode the NWA aligned words of one line (+ look-ahead): returns non-zero iff a sequence byte is not A/C/G/T;
// PK receives the 2-bit codes LSB-first (base j of the line at bits 2(j%16) of PK[j/16]; the '\n' is skipped).
// GAP = bytes between the line's last base and the look-ahead bases: 1 (the '\n' of wrapped FASTA) or 0 (virtual lines cut
// out of one long line, see vl_process_piece).
// PACK = false: the validity test only (PK untouched) -- the line kernel packs while it counts (ln_pack_count_pairs).
template <int LW, int GAP = 1, bool PACK = true>
__device__ __forceinline__ uint32_t ln_decode(const uint32_t (&x)[LineGeom<LW>::NWA + 1], uint32_t (&PK)[LineGeom<LW>::NPK + 1]) {
    using G = LineGeom<LW>;
    static_assert(GAP == 0 || GAP == 1, "gap");
    constexpr uint32_t MPACK = (1u << 23) + (1u << 17) + (1u << 11) + (1u << 5);   // byte3 = c0 | c1<<2 | c2<<4 | c3<<6
    uint32_t acc4 = 0, accC = 0;
    uint32_t pk[G::NWA];
#pragma unroll
    for (int i = 0; i < G::NWA; i++) {
        const uint32_t w = x[i];
        pk[i] = PACK ? (w & 0x06060606u) * MPACK : 0u;
        // at bit 4 of every byte: bit4 == (bit2 & ~bit1)  [T is the only base with bit 4]  and  bit0 != bit4
        const uint32_t b = w * 4u, c = w * 8u, d = w * 16u;
        const uint32_t X = w ^ (b & ~c);
        const uint32_t Y = X | ~(w ^ d);
        const uint32_t Z = w ^ 0x40404040u;   // bits 7,6,3 must read 0,1,0
        uint32_t m = 0;                       // bytes of this word that are sequence: [0,LW) and (LW, LW+LA]
#pragma unroll
        for (int bb = 0; bb < 4; bb++) {
            const int byte = 4 * i + bb;
            if (byte < LW || (byte >= LW + GAP && byte < LW + GAP + G::LA)) m |= 0xFFu << (8 * bb);
        }
        if (m == 0xFFFFFFFFu) { acc4 |= Y; accC |= Z; }
        else if (m != 0) { acc4 |= Y & m; accC |= Z & m; }
    }
    if (!PACK) return (acc4 & 0x10101010u) | (accC & 0xC8C8C8C8u);
    constexpr int NFULL = (LW / 4) / 4;   // groups of four full words -> one register, three byte permutes
#pragma unroll
    for (int gq = 0; gq < NFULL; gq++) {
        const uint32_t r1 = __byte_perm(pk[4 * gq], pk[4 * gq + 1], 0x0073);       // [.., .., p1.b3, p0.b3]
        const uint32_t r2 = __byte_perm(pk[4 * gq + 2], pk[4 * gq + 3], 0x0073);   // [.., .., p3.b3, p2.b3]
        PK[gq] = __byte_perm(r1, r2, 0x5410);
    }
#pragma unroll
    for (int i = NFULL; i <= G::NPK; i++) PK[i] = 0;
#pragma unroll
    for (int i = 4 * NFULL; i < G::NWA; i++) {
        const uint32_t g8 = pk[i] >> 24;   // 4 codes, byte 0's in bits 1:0
#pragma unroll
        for (int seg = 0; seg < 2; seg++) {
            int b0 = (seg == 0) ? 0 : LW + GAP - 4 * i;                // line bases | look-ahead bases
            int b1 = (seg == 0) ? LW - 4 * i : LW + GAP + G::LA - 4 * i;
            b0 = b0 < 0 ? 0 : b0;
            b1 = b1 > 4 ? 4 : b1;
            if (b1 > b0) {
                const int n = b1 - b0;
                const int byte = 4 * i + b0;
                const int pos = byte < LW ? byte : byte - GAP;        // sequence bytes before this one
                const uint32_t val = (g8 >> (2 * b0)) & ((1u << (2 * n)) - 1u);
                const int sh = 2 * (pos & 15);
                PK[pos >> 4] |= val << sh;
                if (sh + 2 * n > 32) PK[(pos >> 4) + 1] |= val >> (32 - sh);
            }
        }
    }
    return (acc4 & 0x10101010u) | (accC & 0xC8C8C8C8u);
}

// Count the NPAIR 8-mers ("pairs" of 7-mers) of one decoded line into the pair histogram.  The 8-mer v (16 bits, first base
// in bits 1:0) is cut out on bits 16:1 of sv: word = v >> 1 (byte address sv & 0x1FFFC); v's lowest bit picks the addend,
// 1 (low half: every pair of the word) or 0x20001 (the high half counts the odd pairs, TWICE) -- and that bit is found on
// bit 17 of the register cut out for the pair four places earlier, so a pair costs three ALU instructions and the RED:
// funnel shift, two LOP3 (`one` is an opaque register holding 1).  Why word = v >> 1 and not v & 0x7FFF: the bank of a word
// is its low five index bits; the low bit of a base code says G/C, and on a genome that is not 50 % GC a bank index with
// three such bits (v bits 4:0) costs 3.9 wavefronts per RED against 3.8 with two (v bits 5:1) -- measured, and the kernel's
// time follows the shared-memory wavefronts.
template <int LW>
__device__ __forceinline__ uint32_t ln_stream32(const uint32_t (&PK)[LineGeom<LW>::NPK + 1], int s) {   // stream bits s .. s+31
    const int q = s >> 5, r = s & 31;
    return r == 0 ? PK[q] : __funnelshift_r(PK[q], PK[q + 1], r);
}
template <int LW, uint32_t BASE>
__device__ __forceinline__ void ln_count_pairs(const uint32_t (&PK)[LineGeom<LW>::NPK + 1], uint32_t *hist16, uint32_t hbase, uint32_t one) {
    using G = LineGeom<LW>;
    uint32_t sv[G::NPAIR];
#pragma unroll
    for (int i = 0; i < G::NPAIR; i++) sv[i] = i == 0 ? (PK[0] << 1) : ln_stream32<LW>(PK, 4 * i - 1);
#pragma unroll
    for (int i = 0; i < G::NPAIR; i++) {
        const uint32_t par = i >= 4 ? sv[i - 4] : (PK[0] << (17 - 4 * i));   // bit 17 = stream bit 4i = v's lowest bit
        red_shared_add_imm<BASE>(hist16, hbase, sv[i] & 0x1FFFCu, (par & 0x20000u) | one);
    }
}

// The same, packing as it goes: register PK[g] (16 bases) is packed right before the pairs that need it, so the line's
// ~100 packing instructions sit BETWEEN its 40 REDs instead of in front of them.  (With everything packed first the
// REDs come every fourth instruction and the warps pile up behind the shared-memory queue: mio_throttle 3.6 % -> 14 %,
// +1.5 % time although the instruction count fell by a fifth.)  Only for a line already found clean.
template <int LW, int GAP, uint32_t BASE>
__device__ __forceinline__ void ln_pack_count_pairs(const uint32_t (&x)[LineGeom<LW>::NWA + 1], uint32_t *hist16, uint32_t hbase, uint32_t one) {
    using G = LineGeom<LW>;
    constexpr uint32_t MPACK = (1u << 23) + (1u << 17) + (1u << 11) + (1u << 5);
    constexpr int NFULL = (LW / 4) / 4;
    uint32_t PK[G::NPK + 1];
    // the registers past the full groups (line tail + look-ahead) first: they are few and their layout is irregular
#pragma unroll
    for (int i = NFULL; i <= G::NPK; i++) PK[i] = 0;
#pragma unroll
    for (int i = 4 * NFULL; i < G::NWA; i++) {
        const uint32_t g8 = ((x[i] & 0x06060606u) * MPACK) >> 24;
#pragma unroll
        for (int seg = 0; seg < 2; seg++) {
            int b0 = (seg == 0) ? 0 : LW + GAP - 4 * i;
            int b1 = (seg == 0) ? LW - 4 * i : LW + GAP + G::LA - 4 * i;
            b0 = b0 < 0 ? 0 : b0;
            b1 = b1 > 4 ? 4 : b1;
            if (b1 > b0) {
                const int n = b1 - b0;
                const int byte = 4 * i + b0;
                const int pos = byte < LW ? byte : byte - GAP;
                const uint32_t val = (g8 >> (2 * b0)) & ((1u << (2 * n)) - 1u);
                const int sh = 2 * (pos & 15);
                PK[pos >> 4] |= val << sh;
                if (sh + 2 * n > 32) PK[(pos >> 4) + 1] |= val >> (32 - sh);
            }
        }
    }
    auto pack_group = [&](int gq) {
        const uint32_t p0 = (x[4 * gq] & 0x06060606u) * MPACK, p1 = (x[4 * gq + 1] & 0x06060606u) * MPACK;
        const uint32_t p2 = (x[4 * gq + 2] & 0x06060606u) * MPACK, p3 = (x[4 * gq + 3] & 0x06060606u) * MPACK;
        PK[gq] = __byte_perm(__byte_perm(p0, p1, 0x0073), __byte_perm(p2, p3, 0x0073), 0x5410);
    };
    uint32_t sv[G::NPAIR];
    if (NFULL > 0) pack_group(0);
    int packed = NFULL > 0 ? 1 : 0;   // full groups packed so far (compile-time after unrolling)
#pragma unroll
    for (int i = 0; i < G::NPAIR; i++) {
        // pair i reads stream bits [4i - 1, 4i + 17): registers (4i - 1) >> 5 and, when it crosses, the next one
        const int need = (4 * i + 16) >> 5;      // highest register index it may touch
        if (need >= packed && packed < NFULL) { pack_group(packed); packed++; }
        sv[i] = i == 0 ? (PK[0] << 1) : ln_stream32<LW>(PK, 4 * i - 1);
        const uint32_t par = i >= 4 ? sv[i - 4] : (PK[0] << (17 - 4 * i));
        red_shared_add_imm<BASE>(hist16, hbase, sv[i] & 0x1FFFCu, (par & 0x20000u) | one);
    }
}

// All 32 lanes: count the K-mers that start at stream positions [0, npos) of ONE line whose bytes sit in shared memory.
// Stream position p is byte p for p < skip and byte p + 1 from there on (skip = index of the one '\n' that is not part
// of the stream; pass a large value when there is none); nstream positions exist.  A K-mer counts iff its K bytes are
// all A/C/G/T, i.e. every other byte is taken as a window break -- the callers make sure no '\n' is among them.
template <int K, class Sink>
__device__ __forceinline__ void ln_coop_line(const uint8_t *bytes, int npos, int nstream, int skip, Sink sink) {
    const int lane = threadIdx.x & 31;
    for (int pos = lane; pos < npos; pos += 32) {
        bool valid = pos + K <= nstream;
        uint32_t kmer = 0;
        if (valid) {
#pragma unroll
            for (int t = 0; t < K; t++) {
                const int sp = pos + t;
                const uint32_t c = bytes[sp < skip ? sp : sp + 1];
                valid = valid && is_base(c);
                kmer = (kmer << 2) | ((c >> 1) & 3u);
            }
        }
        if (valid) sink(kmer << 2);
    }
}

#ifdef KF_PIECE_TIMING
// developer instrumentation (tools/ubench only): SM-clock sums over all warps / CTAs
// [0] warp cycles claiming+processing units  [1] warp cycles waiting at the end-of-piece barrier  [2] CTA cycles in the flush
// [3] pieces  [4] CTA cycles total  [5] CTA cycles anchor search + cursor reset  [6] warp cycles in the fold loop
// ([1] = arrival at the flush until the checksum is known: barrier wait + low-half sums; [2] = zeroing + last barrier)
__device__ unsigned long long g_piece_timing[16];
// [8] clean windows  [9] their warp cycles (decode + count)  [10] windows with a dirty lane  [11] their warp cycles
// [12] grid breaks  [13] warp cycles in the exact generic region  [14] warp cycles finding the next line start at a break
#define KF_T(var) const long long var = clock64()
#define KF_TADD(i, v) do { if ((threadIdx.x & 31) == 0) atomicAdd(&g_piece_timing[i], (unsigned long long)(v)); } while (0)
#else
#define KF_T(var)
#define KF_TADD(i, v)
#endif

struct LnUnit {      // a claimed byte range of the current file piece
    uint64_t Us, Ue;   // lines that start in [Us, Ue) are this warp's
};

// One warp: claim units of the piece [A, Xe) of file [F0, F1) until none is left.  A is a line start.
// s_cursor: shared byte cursor relative to A.  Returns the number of pairs issued through npairs.
template <int LW, uint32_t BASE>
__device__ __forceinline__ void ln_process_piece(const uint8_t *__restrict__ arena, const uint64_t A, const uint64_t Xe,
                                                 const uint64_t F0, const uint64_t F1, uint8_t *buf, uint32_t *wscr,
                                                 uint64_t *bar, uint32_t &par, uint32_t *s_cursor, const uint32_t nwarps, uint32_t *hist16,
                                                 SingleSink gs, uint32_t &npairs) {
    using G = LineGeom<LW>;
    constexpr int K = 7;
    const int lane = threadIdx.x & 31;
    const uint32_t hbase = smem_addr(hist16);
    const uint32_t one = opaque_one();
    const GlobalSrc gsrc{arena};
    auto emit = [&](uint32_t xk) { gs(xk << 2); };
    if (A >= Xe) return;
#ifdef KF_PIECE_TIMING
    unsigned long long tacc[7] = {0, 0, 0, 0, 0, 0, 0};   // slots 8..14 of g_piece_timing, added once per piece
#define KF_LADD(i, v) tacc[(i) - 8] += (unsigned long long)(v)
#else
#define KF_LADD(i, v)
#endif
    const uint64_t plen = Xe - A;
    // guided self-scheduling: a 1/(2 * nwarps) share of what is left, in whole windows, between 2 and 64 windows.
    // The piece's last TAIL bytes are handed out FIRST (cursor range [0, TAIL)): the end of a file -- short last line,
    // missing newline -- takes the slow exact path, which must not be what the other warps wait for at the barrier.
    const uint64_t TAIL = plen > 4ull * G::WIN ? 2ull * G::WIN : 0ull;
    const uint64_t body = plen - TAIL;   // cursor range [TAIL, TAIL + body) is the piece's [0, body)
    auto claim = [&](LnUnit &U) -> bool {
        uint32_t old = 0, sz = 0;
        if (lane == 0) {
            const uint64_t seen = *reinterpret_cast<volatile uint32_t *>(s_cursor);
            if (TAIL && seen == 0) sz = (uint32_t)TAIL;
            else {
                const uint64_t left = seen < plen ? plen - seen : 0;
                uint64_t nwin = left / ((uint64_t)G::WIN * 2u * nwarps);
                nwin = nwin < 2 ? 2 : (nwin > 64 ? 64 : nwin);
#ifdef KF_EMU_RANDOM_UNITS
                { extern unsigned g_emu_seed; uint64_t z = (seen + g_emu_seed) * 0x9E3779B97F4A7C15ull; z ^= z >> 29; nwin = 1 + (z % 9); }   // tests: any unit size must do
#endif
                sz = (uint32_t)(nwin * G::WIN);
            }
            old = atomicAdd(s_cursor, sz);
        }
        old = __shfl_sync(FULL, old, 0);
        sz = __shfl_sync(FULL, sz, 0);
        if ((uint64_t)old >= plen) return false;
        if ((uint64_t)old < TAIL) {   // only the very first claim (old == 0, made while the cursor read 0, so sz == TAIL)
            U.Us = A + body;
            U.Ue = A + plen;
            return true;
        }
        U.Us = A + (old - TAIL);
        U.Ue = U.Us + sz;
        if (U.Ue > A + body) U.Ue = A + body;
        return true;
    };
    LnUnit U;
    bool have = claim(U);
    // staging of the window whose bytes start at `from` (a multiple of 16): one bulk copy, or two halves of 16 lines
    auto issue = [&](uint64_t from) {
        if constexpr (G::DUAL) stage_issue2(buf, arena + from, buf + G::HOFF, arena + from + 16 * G::P, G::HCOPY, bar);
        else stage_issue(buf, arena + from, G::STAGE, bar);
    };
    // offset of line l of the window in the staging buffer, less woff
    const uint32_t lane_off = G::DUAL ? (uint32_t)(lane & 15) * G::P + (uint32_t)(lane >> 4) * G::HOFF : (uint32_t)lane * G::P;
    auto line_off = [&](uint32_t l) -> uint32_t { return G::DUAL ? (l & 15u) * G::P + (l >> 4) * G::HOFF : l * G::P; };
    // window state: B = first line start of the window (unknown while `search`), woff = B - staged base
    uint64_t B = 0, base = 0;
    uint32_t woff = 0;
    bool search = false;
    int strikes = 0;
    // wleft: that many windows from B on are known to lie inside the current unit with a full window after them -- they
    // take the short path below (no break analysis, no unit bookkeeping) as long as all 32 slots are on the grid
    uint32_t wleft = 0;
    auto set_wleft = [&](const LnUnit &u) {
        wleft = (u.Ue > B && u.Ue - B >= 2ull * G::WIN) ? (uint32_t)((u.Ue - B) / (uint64_t)G::WIN) - 1u : 0u;
    };
    auto start_unit = [&](const LnUnit &u) {
        search = u.Us > A;          // A itself is a line start; later units are found from the staged bytes
        base = search ? ((u.Us - 1) & ~15ull) : (u.Us & ~15ull);
        B = u.Us;
        woff = (uint32_t)(B - base);
        strikes = 0;
        wleft = 0;
        if (!search) set_wleft(u);
        KF_SYNCWARP();
        if (lane == 0) issue(base);
    };
    auto refetch = [&](uint64_t at) {   // window whose first line starts at `at` (same unit)
        base = at & ~15ull;
        B = at;
        woff = (uint32_t)(at & 15u);
        search = false;
        set_wleft(U);
        KF_SYNCWARP();
        if (lane == 0) issue(base);
    };
    if (have) start_unit(U);
    while (have) {
        stage_wait(bar, par);
        par ^= 1u;
        if (search) {
            // first line start at or after Us: the byte after the first '\n' in [Us-1, Us+95)
            const uint32_t o0 = (uint32_t)((U.Us - 1) - base);
            int first = 3;
#pragma unroll
            for (int j = 2; j >= 0; j--)
                if (buf[o0 + 3 * lane + j] == 0x0Au) first = j;
            const unsigned Bm = __ballot_sync(FULL, first < 3);
            if (Bm) {
                const int jl = __ffs((int)Bm) - 1;
                const int fj = __shfl_sync(FULL, first, jl);
                B = (U.Us - 1) + 3ull * jl + (uint64_t)fj + 1;
            } else {
                B = fasta_line_start_at_or_after(gsrc, U.Us + 95, F0, F1, lane);
            }
            if (B > F1) B = F1;
            search = false;
            if (B >= U.Ue) {   // no line starts inside this unit
                have = claim(U);
                if (have) start_unit(U);
                continue;
            }
            woff = (uint32_t)(B - base);
            if ((uint64_t)(B - base) > (uint64_t)G::WOFF_MAX) { refetch(B); continue; }
            set_wleft(U);
        }
        // ---- pull my line (+ look-ahead) into registers, byte-aligned ----
        const uint32_t o = woff + lane_off;
        const uint32_t *sw = reinterpret_cast<const uint32_t *>(buf) + (o >> 2);
        const uint32_t ash = (o & 3u) * 8u;
        uint32_t x[G::NWA + 1];
#pragma unroll
        for (int i = 0; i <= G::NWA; i++) x[i] = sw[i];
#pragma unroll
        for (int i = 0; i < G::NWA; i++) x[i] = __funnelshift_r(x[i], x[i + 1], ash);
        const bool nl_byte = ((x[LW / 4] >> (8 * (LW & 3))) & 0xFFu) == 0x0Au;   // my slot ends with a '\n'
        const uint64_t curB = B;
        uint32_t f = 32;
        uint64_t gen_lo = 0, gen_hi = 0;   // exact generic region (empty when the window was clean)
        if (wleft != 0 && __ballot_sync(FULL, nl_byte) == FULL) {
            // ---- the common case: 32 slots on the grid, this window and the next one inside the unit (and so inside the
            // file): the next window follows this one ----
            wleft--;
            strikes = 0;
            B += (uint64_t)G::WIN;
            base = B & ~15ull;
            woff = (uint32_t)B & 15u;
            KF_SYNCWARP();
            if (lane == 0) issue(base);
        } else {
        // on the grid: my slot ends with a '\n' AND lies inside the file (a short last line + arena padding + the next
        // file's header can add up to exactly one slot -- profiles/r01: one k-mer in 5e9 counted across two files)
        const bool nl_ok = nl_byte && B + (uint64_t)(lane + 1) * G::P <= F1;
        // ---- who is on the grid: lines that start before the unit's end ----
        const uint64_t rem = U.Ue - B;                                  // > 0
        const uint32_t nact = rem >= (uint64_t)G::WIN ? 32u : (uint32_t)((rem + G::P - 1) / G::P);
        const bool active = (uint32_t)lane < nact;
        const unsigned bad = __ballot_sync(FULL, active && !nl_ok);
        f = bad ? (uint32_t)(__ffs((int)bad) - 1) : nact;
        // ---- where the next window is; get its copy going before the decode ----
        const uint64_t sf = B + (uint64_t)f * G::P;     // first byte not covered by lanes [0, f)
        uint64_t q = sf;                                // next line start to process on the grid
        bool fast_break = false;
        if (f < nact) {
            // the line at sf breaks the grid.  Usual case: the short last line of a record, followed by a header line or
            // the end of the file -- then its k-mers end with it and all lanes count them together straight from the
            // staged window.  Anything else ([sf, q) holds another kind of line) goes to the exact generic path below.
            // q = the next line start that is not a header: where the grid restarts.
            KF_T(tq0);
            const uint8_t *lb = buf + woff + line_off(f);
            const int lim = F1 - sf < (uint64_t)G::P ? (int)(F1 - sf) : G::P;   // bytes of the slot that belong to the file
            int nlp = G::P + 1;   // index of the line's '\n' (none within LW + 1 bytes: the line is longer than the grid's)
#pragma unroll
            for (int i = 0; i < (G::P + 31) / 32; i++) {
                const int bi = lane + 32 * i;
                const unsigned m = __ballot_sync(FULL, bi < lim && lb[bi] == 0x0Au);
                if (m && nlp > G::P) nlp = 32 * i + __ffs((int)m) - 1;
            }
            if (nlp > G::P && lim < G::P) nlp = lim;   // the file ends inside the slot without a final '\n'
            // (two halves: the one that holds line f)
            const WindowSrc wsrc = G::DUAL ? WindowSrc{arena, buf + (f >> 4) * G::HOFF, base + (uint64_t)(f >> 4) * (16 * G::P), (uint32_t)G::HCOPY}
                                           : WindowSrc{arena, buf, base, (uint32_t)G::STAGE};
            const bool is_hdr = lb[0] == (uint8_t)'>';   // sf is a line start: a '>' here opens a header line
            if (nlp <= G::P && is_hdr) {
                fast_break = true;                        // a header line that ends inside the slot: nothing to count
                q = sf + (uint64_t)nlp + 1 < F1 ? sf + (uint64_t)nlp + 1 : F1;
            } else if (nlp < LW && !is_hdr) {
                const uint64_t after = sf + (uint64_t)nlp + 1;
                fast_break = after >= F1 || lb[nlp + 1] == (uint8_t)'>';
                q = after < F1 ? after : F1;
                if (fast_break) ln_coop_line<K>(lb, nlp, nlp, 1 << 20, gs);
            } else {
                q = fasta_line_start_at_or_after(wsrc, sf + 1, F0, F1, lane);
            }
            while (q < U.Ue && wsrc.byte(q) == (uint32_t)'>') q = fasta_line_start_at_or_after(wsrc, q + 1, F0, F1, lane);
            strikes = (f == 0 && !fast_break) ? strikes + 1 : 0;
            KF_T(tq1);
            KF_LADD(14, tq1 - tq0);
        } else {
            strikes = 0;
        }
        gen_lo = sf;
        gen_hi = fast_break ? sf : q;
        if (q < U.Ue && strikes >= 4) {
            // this stretch is not on the grid (another width, blank lines ...): finish the unit generically
            gen_hi = fasta_line_start_at_or_after(gsrc, U.Ue, F0, F1, lane);
            q = U.Ue;
        }
        if (q < U.Ue) refetch(q);
        else {
            have = claim(U);   // (nothing below looks at the unit that has just been finished)
            if (have) start_unit(U);
        }
        }
        // ---- decode + count the lanes on the grid ----
        KF_T(td0);
#ifdef KF_PIECE_TIMING
        bool any_dirty = false;
#endif
        bool dirty = false;
        if ((uint32_t)lane < f) {
            uint32_t PK[G::NPK + 1];
            const uint32_t anyV = ln_decode<LW, 1, false>(x, PK);
            if (anyV == 0) {
                ln_pack_count_pairs<LW, 1, BASE>(x, hist16, hbase, one);
                npairs += G::NPAIR;
            } else {
                dirty = true;
            }
        }
        // lines that hold a non-ACGT byte (or whose look-ahead does): the window buffer is already being refilled, so the
        // line is fetched again from global memory (L2) into the warp's scratch, one word per lane, and all lanes count
        // its k-mers together.  A '\n' among its bytes (a line shorter than the grid's inside the slot, or a next line
        // shorter than the look-ahead) is not a window break: such a line goes to the exact byte walker.
        unsigned D = __ballot_sync(FULL, dirty);
#ifdef KF_PIECE_TIMING
        any_dirty = D != 0;
#endif
        while (D) {
            const int j = __ffs((int)D) - 1;
            D &= D - 1;
            const uint64_t sl = curB + (uint64_t)j * G::P;
            const uint64_t wb = sl & ~3ull;
            KF_SYNCWARP();
            for (int i = lane; i < LN_SCR_WORDS; i += 32) wscr[i] = __ldg(reinterpret_cast<const uint32_t *>(arena + wb) + i);
            KF_SYNCWARP();
            const uint8_t *lb = reinterpret_cast<const uint8_t *>(wscr) + (uint32_t)(sl & 3u);
            bool has_nl = false;
#pragma unroll
            for (int i = 0; i < (G::P + G::LA + 31) / 32; i++) {
                const int bi = lane + 32 * i;
                has_nl = has_nl || (bi < G::P + G::LA && bi != LW && lb[bi] == 0x0Au);
            }
            if (__ballot_sync(FULL, has_nl) != 0) {
                if (lane == j) fasta_walk_lane<K>(gsrc, sl, sl + G::P, false, true, emit);
            } else if (lb[0] != (uint8_t)'>') {   // (a header line exactly as long as a sequence line holds no k-mers)
                ln_coop_line<K>(lb, LW, LW + G::LA, LW, gs);
            }
        }
#ifdef KF_PIECE_TIMING
        {
            const bool wd = any_dirty;
            const long long td1 = clock64();
            KF_LADD(wd ? 10 : 8, 1);
            KF_LADD(wd ? 11 : 9, td1 - td0);
        }
#endif
        KF_T(tg0);
        if (gen_lo < gen_hi) lg_generic_region<K>(gsrc, gen_lo, gen_hi, (uint32_t)(F0 / CHUNK), gs);
#ifdef KF_PIECE_TIMING
        if (gen_lo < gen_hi) { const long long tg1 = clock64(); KF_LADD(12, 1); KF_LADD(13, tg1 - tg0); }
#endif
    }
#ifdef KF_PIECE_TIMING
    for (int i = 0; i < 7; i++) if (tacc[i]) KF_TADD(8 + i, tacc[i]);
#endif
}


// byte-mask helpers (also used by the FASTQ kernels below)
__device__ __forceinline__ uint32_t nonzero_bytes(uint32_t v) {   // 0x80 in every byte of v that is non-zero
    return (((v & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | v) & 0x80808080u;
}
__device__ __forceinline__ uint32_t movemask4(uint32_t f) {       // f has 0x80 flags; -> 4-bit mask, byte 0 in bit 0
    return ((f >> 7) * 0x01020408u) >> 24;
}
__device__ __forceinline__ uint32_t newline_mask16(const uint4 w) {
    const uint32_t n0 = ~nonzero_bytes(w.x ^ 0x0A0A0A0Au) & 0x80808080u;
    const uint32_t n1 = ~nonzero_bytes(w.y ^ 0x0A0A0A0Au) & 0x80808080u;
    const uint32_t n2 = ~nonzero_bytes(w.z ^ 0x0A0A0A0Au) & 0x80808080u;
    const uint32_t n3 = ~nonzero_bytes(w.w ^ 0x0A0A0A0Au) & 0x80808080u;
    return movemask4(n0) | (movemask4(n1) << 4) | (movemask4(n2) << 8) | (movemask4(n3) << 12);
}

// ================================================================================================
// Virtual lines: long-line ("unwrapped") FASTA through the same pair histogram
// ================================================================================================
// Assemblers write one line per contig (kilobases to megabases).  There is no grid of '\n' to hang a lane on -- but
// inside such a line any 80 bytes are 80 bases, so a warp cuts VIRTUAL lines of VL = 80 bytes wherever it likes: lane l
// takes the bytes [B + 80 l, B + 80 l + 86) of a window that starts at a multiple of 16.  Compared with the wrapped
// kernel: five aligned LDS.128 + one LDS.64 instead of 23 LDS.32 and 22 funnel shifts, no bank conflicts (lane pitch =
// 5 x 16 bytes), nothing to delete.  What has to be known instead is whether a byte range lies in a sequence line or in
// a header line -- and the line's start may be megabytes back:
//   piece  : the state at the piece's first byte comes from an exact backward scan by the whole CTA (only the pieces
//            that begin inside a file need one).
//   unit   : a unit (byte range [Us, Ue), Us = A + n * 2560, claimed from the shared cursor) owns the k-mers whose first
//            base lies in it.  Its state at Us: the last '\n' of the 256 bytes before Us decides; if there is none the
//            unit ASSUMES "sequence line" and notes Us in a shared list.
//   breaks : a lane whose 86 bytes hold a '\n' (or reach the file's end) stops the grid; the warp walks the line
//            structure from there (tail of the line, header lines -- each one noted with its first and last byte --,
//            blank lines) until a sequence line is long enough to carry the grid again, which restarts on the next slot
//            boundary of the unit; the part slots on both sides are counted by all lanes together.
//   verify : when the piece is done, an assumed unit start that lies inside a noted header line means text was counted
//            that is no sequence (a header longer than 256 bytes across a unit boundary): the flush then discards the
//            histograms and recounts the piece exactly, like after a wrapped 16-bit half.
#ifdef KF_VL_TIMING
__device__ unsigned long long g_vl_timing[16];   // developer instrumentation: warp cycles [0] walk_structure [1] unit start [2] stage waits [3] windows [4] #walks [5] #units [6] #windows [7] total
__device__ unsigned long long g_vl_cta[4 * 160];   // per CTA: total cycles, pieces, max piece cycles, cycles in the piece-start backward scan
#define VLT(var) const long long var = clock64()
#define VLADD(i, v) do { if ((threadIdx.x & 31) == 0) atomicAdd(&g_vl_timing[i], (unsigned long long)(v)); } while (0)
#else
#define VLT(var)
#define VLADD(i, v)
#endif
constexpr int VL = 80;                       // bytes (= bases) per virtual line
constexpr int VL_WIN = 32 * VL;              // 2,560 bytes per window
constexpr int VL_STAGE = VL_WIN + 16;        // + the last lane's look-ahead
constexpr int VL_LOOKBACK = 256;
constexpr int VL_LOG_SPEC = 1024;            // assumed unit starts per piece (u32 file offsets)
constexpr int VL_LOG_HDR = 512;              // header lines per piece (first byte, terminating '\n': u32 file offsets)

struct VlLog {            // shared memory, per CTA
    uint32_t n_spec, n_hdr, overflow, pad;
    uint32_t spec[VL_LOG_SPEC];
    uint32_t hdr[2 * VL_LOG_HDR];
};

// All 32 lanes: count the K-mers whose first base lies in [lo, hi) of a stretch without '\n' in [lo, hi + K - 1) (the
// caller knows): a K-mer counts iff its K bytes are bases.  Bytes are read where they lie (L1/L2).
template <int K, class Sink>
__device__ __forceinline__ void vl_count_stretch(const uint8_t *__restrict__ arena, uint64_t lo, uint64_t hi, uint64_t limit, Sink sink) {
    if (hi <= lo) return;
    const uint64_t avail = limit - lo;                    // bytes that may be looked at from lo on
    const int nstream = avail > (uint64_t)(1 << 20) ? (1 << 20) : (int)avail;
    ln_coop_line<K>(arena + lo, (int)(hi - lo), nstream, 1 << 30, sink);
}

// One warp, virtual lines: units of the piece [A, Xe) of the file [F0, F1); state_at_A: 0 = A lies inside a sequence line,
// 1 = inside a header line that began before A, 2 = A is a line start.
template <uint32_t BASE>
__device__ __forceinline__ void vl_process_piece(const uint8_t *__restrict__ arena, const uint64_t A, const uint64_t Xe, const uint32_t state_at_A,
                                                 const uint64_t F0, const uint64_t F1, uint8_t *buf, uint32_t *wscr, uint64_t *bar,
                                                 uint32_t &par, uint32_t *s_cursor, const uint32_t nwarps, uint32_t *hist16, SingleSink gs,
                                                 uint32_t &npairs, VlLog *log) {
    constexpr int K = 7;
    constexpr int LW = VL;
    using G = LineGeom<LW>;
    static_assert(G::NWA == 22 && (VL % 16) == 0, "five 16-byte pieces and one 8-byte look-ahead per lane");
    const int lane = threadIdx.x & 31;
    const uint32_t hbase = smem_addr(hist16);
    const uint32_t one = opaque_one();
    const GlobalSrc gsrc{arena};
    auto emit = [&](uint32_t xk) { gs(xk << 2); };
    if (A >= Xe) return;
    const uint64_t plen = Xe - A;
    struct Unit { uint64_t Us, Ue; };
    auto claim = [&](Unit &U) -> bool {
        uint32_t old = 0, sz = 0;
        if (lane == 0) {
            const uint64_t seen = *reinterpret_cast<volatile uint32_t *>(s_cursor);
            const uint64_t left = seen < plen ? plen - seen : 0;
            uint64_t nwin = left / ((uint64_t)VL_WIN * 2u * nwarps);
            nwin = nwin < 2 ? 2 : (nwin > 64 ? 64 : nwin);
#ifdef KF_EMU_RANDOM_UNITS
            { extern unsigned g_emu_seed; uint64_t z = (seen + g_emu_seed) * 0x9E3779B97F4A7C15ull; z ^= z >> 29; nwin = 1 + (z % 5); }
#endif
            sz = (uint32_t)(nwin * VL_WIN);
            old = atomicAdd(s_cursor, sz);
        }
        old = __shfl_sync(FULL, old, 0);
        sz = __shfl_sync(FULL, sz, 0);
        if ((uint64_t)old >= plen) return false;
        U.Us = A + old;
        U.Ue = U.Us + sz < Xe ? U.Us + sz : Xe;
        return true;
    };
    auto log_spec = [&](uint64_t us) {
        if (lane == 0) {
            const uint32_t i = atomicAdd(&log->n_spec, 1u);
            if (i < (uint32_t)VL_LOG_SPEC) log->spec[i] = (uint32_t)(us - F0); else log->overflow = 1u;
        }
    };
    auto log_hdr = [&](uint64_t hs, uint64_t he) {
        if (lane == 0) {
            const uint32_t i = atomicAdd(&log->n_hdr, 1u);
            if (i < (uint32_t)VL_LOG_HDR) { log->hdr[2 * i] = (uint32_t)(hs - F0); log->hdr[2 * i + 1] = (uint32_t)(he - F0); } else log->overflow = 1u;
        }
    };
    // position of the first '\n' at or after q, F1 if there is none (all lanes)
    auto next_nl = [&](uint64_t q) -> uint64_t {
        if (q >= F1) return F1;
        const uint64_t ls = fasta_line_start_at_or_after(gsrc, q + 1, F0, F1, lane);   // byte after the first '\n' >= q
        return (ls == F1 && arena[F1 - 1] != 0x0Au) ? F1 : ls - 1;
    };
    // From the line start `pos`: blank lines and header lines are skipped (headers noted); returns the first byte of the
    // next sequence line, or U.Ue when the unit (or the file) ends first.
    auto skip_headers = [&](const Unit &U, uint64_t pos) -> uint64_t {
        for (;;) {
            if (pos >= F1 || pos >= U.Ue) return U.Ue;
            const uint32_t c = arena[pos];
            if (c == 0x0Au) { pos++; continue; }                 // blank line
            if (c != (uint32_t)'>') return pos;
            const uint64_t he = next_nl(pos);
            log_hdr(pos, he);
            if (he >= F1) return U.Ue;
            pos = he + 1;
        }
    };
    // From `pos` inside a sequence line (the bytes before it are done): count what lies between pos and the place where
    // full virtual lines can go on, walking the line structure; returns that place -- a slot boundary of the unit's
    // grid (U.Us + 80 n) -- or U.Ue.  after_pos: the boundary must lie beyond pos (the unit's last, partial slot).
    auto walk_structure = [&](const Unit &U, uint64_t pos, const bool after_pos) -> uint64_t {
        for (;;) {
#ifdef KF_EMU_TRACE
            if (lane == 0) fprintf(stderr, "  walk pos=%llu Us=%llu Ue=%llu after=%d\n", (unsigned long long)(pos - F0), (unsigned long long)(U.Us - F0), (unsigned long long)(U.Ue - F0), (int)after_pos);
#endif
            if (pos >= U.Ue || pos >= F1) return U.Ue;
            const uint64_t nslot = (pos - U.Us + VL - 1) / VL + ((after_pos && (pos - U.Us) % VL == 0) ? 1 : 0);
            const uint64_t slot = U.Us + nslot * VL;                              // next slot boundary
            const uint64_t hi = slot < U.Ue ? slot : U.Ue;
            // the line's '\n' if it lies before hi + K - 1 (a bounded look: the line may be megabytes long), F1 at the
            // end of the file
            const uint64_t lim = hi + (K - 1) < F1 ? hi + (K - 1) : F1;
            uint64_t p = lim;
            for (uint64_t q0 = pos; q0 < lim && p == lim; q0 += 96) {
                int first = 3;
#pragma unroll
                for (int j = 2; j >= 0; j--) {
                    const uint64_t a = q0 + 3ull * lane + j;
                    if (a < lim && arena[a] == 0x0Au) first = j;
                }
                const unsigned Bm = __ballot_sync(FULL, first < 3);
                if (Bm) {
                    const int jl = __ffs((int)Bm) - 1;
                    p = q0 + 3ull * jl + (uint64_t)__shfl_sync(FULL, first, jl);
                }
            }
            if (p == lim && lim < F1) {
                // the line carries on beyond the boundary (look-ahead included): the grid goes on there
                vl_count_stretch<K>(arena, pos, hi, lim, gs);
                return hi;
            }
            // the line ends first.  What follows its '\n'?  A sequence line (the record goes on: k-mers span the '\n') is
            // the byte walker's case; a header line, the end of the file (or blank lines and then one of them) close it.
            uint64_t nx = p + 1;
            while (nx < F1 && arena[nx] == 0x0Au) nx++;
            const bool closes = nx >= F1 || arena[nx] == (uint8_t)'>';
            if (closes) {
                vl_count_stretch<K>(arena, pos, p < U.Ue ? p : U.Ue, p, gs);
                if (p >= F1) return U.Ue;
                pos = skip_headers(U, p + 1);
            } else {
                if (lane == 0) fasta_walk_lane<K>(gsrc, pos, nx < U.Ue ? nx : U.Ue, false, false, emit);
                KF_SYNCWARP();
                pos = nx;
            }
        }
    };
    Unit U;
    VLT(t_all0);
    bool have = claim(U);
    while (have) {
        VLT(t_u0);
        // ---- where the unit's grid starts: the state at Us decides ----
        uint64_t B;
        uint32_t st;   // 0: inside a sequence line, 1: inside a header line that began before Us, 2: Us is a line start
        if (U.Us == A) {
            st = state_at_A;
        } else {
            const uint64_t lb0 = U.Us - F0 >= (uint64_t)VL_LOOKBACK ? U.Us - VL_LOOKBACK : F0;   // (multiple of 16)
            const uint64_t q = lb0 + 8ull * lane;
            uint2 w = make_uint2(0u, 0u);
            if (q < U.Us) w = __ldg(reinterpret_cast<const uint2 *>(arena + q));
            int last = -1;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const uint32_t c = ((i < 4 ? w.x : w.y) >> (8 * (i & 3))) & 0xFFu;
                if (q + i < U.Us && c == 0x0Au) last = i;
            }
            const unsigned Bm = __ballot_sync(FULL, last >= 0);
            if (Bm) {
                const int j = 31 - __clz((int)Bm);
                const int lj = __shfl_sync(FULL, last, j);
                const uint64_t ls = lb0 + 8ull * j + (uint64_t)lj + 1;   // start of the line that holds Us (<= Us)
                st = ls == U.Us ? 2u : (arena[ls] == (uint8_t)'>' ? 1u : 0u);
            } else if (lb0 == F0) {
                st = 1u;                  // still on the file's first line: the header ('>' is the file's first byte)
            } else {
                st = 0u;                  // assumed: sequence line -- checked when the piece is done
                log_spec(U.Us);
            }
        }
        if (st == 1u) {
            const uint64_t he = next_nl(U.Us);
            if (U.Us == A) log_hdr(A, he);   // (it began before the piece: nobody else notes it)
            B = (he >= F1 || he + 1 >= U.Ue) ? U.Ue : walk_structure(U, skip_headers(U, he + 1), false);
        } else if (st == 2u) {
            B = walk_structure(U, skip_headers(U, U.Us), false);
        } else {
            B = U.Us;
        }
#ifdef KF_EMU_TRACE
        if (lane == 0) fprintf(stderr, "unit Us=%llu Ue=%llu st=%u B=%llu\n", (unsigned long long)(U.Us - F0), (unsigned long long)(U.Ue - F0), st, (unsigned long long)(B - F0));
#endif
        // ---- windows of full virtual lines from B on ----
        VLT(t_u1);
        VLADD(1, t_u1 - t_u0); VLADD(5, 1);
        bool staged = false;
        while (B < U.Ue) {
            const uint32_t nslots = (uint32_t)((U.Ue - B) / VL);
            if (nslots == 0) {   // the piece ends inside a slot
                VLT(t_w0);
                B = walk_structure(U, B, true);
                VLT(t_w1);
                VLADD(0, t_w1 - t_w0); VLADD(4, 1);
                continue;
            }
            const uint32_t nact = nslots < 32u ? nslots : 32u;
            if (!staged) {
                KF_SYNCWARP();
                if (lane == 0) stage_issue(buf, arena + B, VL_STAGE, bar);
            }
            stage_wait(bar, par);
            par ^= 1u;
            staged = false;
            const uint4 *s4 = reinterpret_cast<const uint4 *>(buf + (uint32_t)lane * VL);
            uint32_t x[G::NWA + 1];
#pragma unroll
            for (int i = 0; i < 5; i++) { const uint4 v = s4[i]; x[4 * i] = v.x; x[4 * i + 1] = v.y; x[4 * i + 2] = v.z; x[4 * i + 3] = v.w; }
            { const uint2 v = *reinterpret_cast<const uint2 *>(buf + (uint32_t)lane * VL + VL); x[20] = v.x; x[21] = v.y; }
            x[22] = 0;
            // the next window of this unit, assuming this one is clean: its copy overlaps the decode
            const uint64_t Bn = B + (uint64_t)VL_WIN;
            if (nact == 32u && Bn + VL <= U.Ue) {
                KF_SYNCWARP();
                if (lane == 0) stage_issue(buf, arena + Bn, VL_STAGE, bar);
                staged = true;
            }
            uint32_t PK[G::NPK + 1];
            const bool active = (uint32_t)lane < nact;
            const uint32_t anyV = ln_decode<LW, 0, false>(x, PK);
            const bool dirty = active && anyV != 0;
            const unsigned D = __ballot_sync(FULL, dirty);
            uint32_t f = nact;   // lanes [0, f) are virtual lines of one sequence line
            if (D) {
                // which dirty lanes hold a '\n' or reach beyond the file?  (the raw bytes are still in x)
                bool brk = false;
                if (dirty) {
                    const uint64_t sl = B + (uint64_t)lane * VL;
                    brk = sl + VL + (K - 1) > F1;
#pragma unroll
                    for (int i = 0; i < G::NWA; i++) {
                        const uint32_t z = x[i] ^ 0x0A0A0A0Au;
                        uint32_t m = (z - 0x01010101u) & ~z & 0x80808080u;
                        if (i == G::NWA - 1) m &= 0x00008080u;   // (bytes 86, 87 are not mine)
                        brk = brk || m != 0;
                    }
                }
                const unsigned Bk = __ballot_sync(FULL, brk);
                if (Bk) f = (uint32_t)(__ffs((int)Bk) - 1);
            }
            if ((uint32_t)lane < f && active && !dirty) {
                ln_pack_count_pairs<LW, 0, BASE>(x, hist16, hbase, one);
                npairs += G::NPAIR;
            }
            // dirty lanes before the break: N runs, IUPAC codes ... -- all lanes count such a line together
            unsigned Dn = D & (f >= 32u ? 0xFFFFFFFFu : ((1u << f) - 1u));
            while (Dn) {
                const int j = __ffs((int)Dn) - 1;
                Dn &= Dn - 1;
                const uint64_t sl = B + (uint64_t)j * VL;
                vl_count_stretch<K>(arena, sl, sl + VL, sl + VL + (K - 1), gs);
            }
            if (f < nact) {
                // the grid breaks in slot f: drop the copy that was started for the next window, walk the structure
                VLT(t_w0);
                if (staged) { stage_wait(bar, par); par ^= 1u; staged = false; }
                B = walk_structure(U, B + (uint64_t)f * VL, true);   // (its '\n' lies within the slot's 86 bytes: go beyond it)
                VLT(t_w1);
                VLADD(0, t_w1 - t_w0); VLADD(4, 1);
            } else {
                B += (uint64_t)nact * VL;
            }
        }
        have = claim(U);
    }
    VLT(t_all1);
    VLADD(7, t_all1 - t_all0);
    (void)wscr;
}

// BASE: shared-window address of the pair histogram (= start of the dynamic shared memory: the kernel has no static
// shared memory) as a compile-time constant, or 0 when it is not known -- the host asks kf_smem_base_probe_kernel once.
// VIRT: the files of this launch are long-line FASTA (file_P == KF_P_VIRTUAL), counted through virtual lines of 80 bytes
// (vl_process_piece); LW must be 80 then (same shared-memory layout).
constexpr uint32_t KF_P_VIRTUAL = 0xFFFFu;
// the wrapped widths that have a line-kernel instantiation, and where the batch's count of such files is kept
__host__ __device__ constexpr int line_width_slot(int lw) {   // index into width_counts (0 = generic kernel, 4 = long lines)
    return lw == 60 ? 1 : lw == 70 ? 2 : lw == 80 ? 3 : lw == 100 ? 5 : lw == 50 ? 7 : 0;
}
// dynamic shared memory of count_fasta_lines_kernel: pair + singles histograms, one staging buffer and one barrier per
// warp, per-warp scratch and partial sums, cursor words, the backward scan's result, the virtual-line log
template <int LW, bool VIRT = false> constexpr size_t lines_kernel_smem(int nwarps) {
    return (32768 + 8192) * sizeof(uint32_t) + (size_t)nwarps * LineGeom<LW>::STAGE + (size_t)nwarps * sizeof(uint64_t) +
           ((LN_SCR_WORDS + 3) * (size_t)nwarps + 4) * sizeof(uint32_t) + 16 + (VIRT ? sizeof(VlLog) : 0);
}
template <int LW, int THREADS, uint32_t BASE, bool VIRT = false>
__global__ void __launch_bounds__(THREADS, 1)
count_fasta_lines_kernel(const uint8_t *__restrict__ arena, const Tile *__restrict__ tiles, const int *__restrict__ cta_begin,
                         const uint32_t *__restrict__ file_P, const uint64_t *__restrict__ file_off,
                         const uint64_t *__restrict__ file_len, unsigned long long *__restrict__ g_fwd,
                         const uint32_t *__restrict__ file_row, const uint32_t *__restrict__ cta_first_rank, int cta_stride,
                         const uint32_t *__restrict__ width_counts) {
    // LW == 0: every supported kind of file in ONE launch (the piece's width picks the code): no empty launches, and a
    // batch of mixed widths keeps all CTAs busy.  Staging is sized for the widest line then.
    constexpr bool ALL = LW == 0;
    using G = LineGeom<ALL ? LN_MAX_LW : LW>;
    static_assert(!(VIRT || ALL) || VL_STAGE <= G::STAGE, "virtual lines fit the staging buffers");
    static_assert(!VIRT || LW == 80, "virtual lines use the 80-column layout");
    constexpr uint32_t MYP = VIRT ? KF_P_VIRTUAL : (uint32_t)G::P;
    static_assert(ALL || VIRT || line_width_slot(LW) != 0, "no slot for this width");
    if (ALL) {
        uint32_t any = 0;
#pragma unroll
        for (int i = 1; i < 8; i++) any |= width_counts[i];
        if (any == 0) return;   // every file goes to the generic kernel (uniform exit)
    } else if (width_counts[VIRT ? 4 : line_width_slot(LW)] == 0) return;   // no file of this kind in the batch (uniform exit)
    constexpr int NWARPS = THREADS / 32;
    constexpr int NWORDS = 32768;
    constexpr int NB7 = 16384;
    KF_DYN_SMEM(uint32_t, smem);
    constexpr int NSWORDS = NB7 / 2;
#ifndef KF_EMU
    if (BASE != 0 && smem_addr(smem) != BASE) __trap();   // (the immediate-address REDs would land elsewhere: fail loudly)
#endif
    uint32_t *hist16 = smem;               // pairs: 65,536 8-mer bins, two u16 halves per word
    uint32_t *single16 = smem + NWORDS;    // rare paths: 16,384 7-mer bins, two u16 halves per word
    uint8_t *stage_base = reinterpret_cast<uint8_t *>(single16 + NSWORDS);
    uint64_t *bars = reinterpret_cast<uint64_t *>(stage_base + (size_t)NWARPS * G::STAGE);
    uint32_t *s_wscr = reinterpret_cast<uint32_t *>(bars + NWARPS);   // per warp: LN_SCR_WORDS words, one dirty line re-fetched from global
    uint32_t *s_part = s_wscr + LN_SCR_WORDS * NWARPS;                          // per warp: pairs issued | pair low-half sums | singles half sums
    uint32_t *s_nsingle = s_part + 3 * NWARPS;                        // singles issued
    uint32_t *s_cursor = s_nsingle + 1;
    unsigned long long *s_found = reinterpret_cast<unsigned long long *>((reinterpret_cast<uintptr_t>(s_cursor + 3) + 7) & ~(uintptr_t)7);
    VlLog *vlog = reinterpret_cast<VlLog *>(s_found + 1);                                   // VIRT only
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < NWORDS + NSWORDS; i += THREADS) smem[i] = 0;
    if (threadIdx.x == 0) { *s_nsingle = 0; *s_cursor = 0; s_cursor[1] = 0; }
    if ((VIRT || ALL) && threadIdx.x == 0) { vlog->n_spec = 0; vlog->n_hdr = 0; vlog->overflow = 0; }
    uint8_t *buf = stage_base + (size_t)warp * G::STAGE;
    uint64_t *bar = bars + warp;
    uint32_t par = 0;
    if (lane == 0) stage_bar_init(bar);
    __syncthreads();
    uint32_t npairs = 0;
    int cur_file = -1;
    uint64_t file_lo = 0, file_hi = 0;   // byte range of cur_file this CTA has processed (for the exact recount)
    const int tb0 = cta_begin[blockIdx.x * cta_stride], tb1 = cta_begin[(blockIdx.x + 1) * cta_stride];
    const int first_file = tb0 < tb1 ? (int)tiles[tb0].file : -1;

    auto flush = [&](int file) {
        // checksums (a u16 half may have wrapped): pairs issued vs the sum of the pair histogram's low halves, singles
        // issued vs the sum of all halves of the singles histogram -- through per-warp partial sums
        uint32_t np = npairs;
        npairs = 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) np += __shfl_xor_sync(FULL, np, o);
        if (lane == 0) s_part[warp] = np;
        KF_T(tw0);
        __syncthreads();
        const bool virt = VIRT || (ALL && file_P[file] == KF_P_VIRTUAL);
        if (virt) {
            // an assumed unit start inside a noted header line: text was counted that is no sequence
            const uint32_t ns = vlog->n_spec < (uint32_t)VL_LOG_SPEC ? vlog->n_spec : (uint32_t)VL_LOG_SPEC;
            const uint32_t nh = vlog->n_hdr < (uint32_t)VL_LOG_HDR ? vlog->n_hdr : (uint32_t)VL_LOG_HDR;
            bool bad = threadIdx.x == 0 && vlog->overflow != 0;
            for (uint32_t i = threadIdx.x; i < ns; i += THREADS) {
                const uint32_t us = vlog->spec[i];
                for (uint32_t h = 0; h < nh; h++) bad = bad || (vlog->hdr[2 * h] < us && us <= vlog->hdr[2 * h + 1]);
            }
            if (bad) s_cursor[1] = 1u;   // (read after the next barrier)
        }
        uint32_t low = 0, sing = 0, big = 0;
        for (int i = threadIdx.x; i < NWORDS; i += THREADS) { const uint32_t v = hist16[i]; low += v & 0xFFFFu; big |= v; }
        for (int i = threadIdx.x; i < NSWORDS; i += THREADS) { const uint32_t v = single16[i]; sing += (v & 0xFFFFu) + (v >> 16); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { low += __shfl_xor_sync(FULL, low, o); sing += __shfl_xor_sync(FULL, sing, o); big |= __shfl_xor_sync(FULL, big, o); }
        // a word whose low half reached 32,768 may have a wrapped high half (it counts twice): reported through the
        // checksum by making it fail
        if (lane == 0) { s_part[NWARPS + warp] = (big & 0x8000u) ? 0xFFFFFFFFu : low; s_part[2 * NWARPS + warp] = sing; }
        __syncthreads();
        unsigned long long tp = 0, tl = 0, ts = 0;
        bool half_ok = true;
#pragma unroll
        for (int w = 0; w < NWARPS; w++) { tp += s_part[w]; tl += s_part[NWARPS + w]; ts += s_part[2 * NWARPS + w]; half_ok = half_ok && s_part[NWARPS + w] != 0xFFFFFFFFu; }
        const bool ok = half_ok && tp == tl && ts == (unsigned long long)*s_nsingle && !(virt && s_cursor[1] != 0);
        // every CTA that holds a piece of the file owns one row of it (file_row[file] + its rank among those CTAs; the
        // fold kernel sums the rows), so the row is WRITTEN, every bin, with plain 16-byte stores: no global atomics.
        // Only the first file of a CTA's tile range can have begun in an earlier CTA: its rank comes from the host.
        const uint32_t rank = (file == first_file) ? cta_first_rank[blockIdx.x] : 0u;
        unsigned long long *g = g_fwd + ((size_t)file_row[file] + rank) * NB7;
#ifdef KF_PIECE_TIMING
        long long tw1 = clock64();
        if (tp == 0xFFFFFFFFFFFFull) tw1 = 0;   // depends on the shared reads above: taken after the barrier released
        KF_TADD(1, tw1 - tw0);
#endif
        // xk: 7-mer with its digits reversed (first base in bits 1:0) -- the orientation of both histograms and of this
        // file's rows in g_fwd (the fold kernel undoes it).  xk is the first 7-mer of the 8-mers xk + a*16384 (word
        // a*8192 + (xk >> 1); odd xk: half of the high half, even xk: low - high / 2) and the second 7-mer of the 8-mers
        // 4xk .. 4xk+3 (words 2xk, 2xk+1: sum of the low halves).  One thread: xk = 2j and 2j+1.
#pragma unroll 4
        for (int j = threadIdx.x; j < NB7 / 2; j += THREADS) {
            unsigned long long c0 = 0, c1 = 0;
            if (ok) {
                const uint4 pw = *reinterpret_cast<const uint4 *>(hist16 + 4 * j);
                const uint32_t sw = single16[j];
                c0 = (pw.x & 0xFFFFu) + (pw.y & 0xFFFFu) + (sw & 0xFFFFu);
                c1 = (pw.z & 0xFFFFu) + (pw.w & 0xFFFFu) + (sw >> 16);
#pragma unroll
                for (int a = 0; a < 4; a++) {
                    const uint32_t w = hist16[a * 8192 + j];
                    c0 += (w & 0xFFFFu) - (w >> 17);
                    c1 += w >> 17;
                }
            }
            reinterpret_cast<ulonglong2 *>(g)[j] = make_ulonglong2(c0, c1);
        }
#ifdef KF_PIECE_TIMING
        const long long tw2 = clock64();
        KF_TADD(6, tw2 - tw1);
#endif
        __syncthreads();
        uint4 *h4 = reinterpret_cast<uint4 *>(smem);
        for (int i = threadIdx.x; i < (NWORDS + NSWORDS) / 4; i += THREADS) h4[i] = make_uint4(0, 0, 0, 0);
        if (threadIdx.x == 0) { *s_nsingle = 0; s_cursor[1] = 0; }
        if ((VIRT || ALL) && threadIdx.x == 0) { vlog->n_spec = 0; vlog->n_hdr = 0; vlog->overflow = 0; }
        __syncthreads();
#ifdef KF_PIECE_TIMING
        { long long tw3 = clock64(); if (*s_nsingle == 12345u) tw3 = 0; KF_TADD(2, tw3 - tw2); if (threadIdx.x == 0) atomicAdd(&g_piece_timing[3], 1ull); }
#endif
#ifdef KF_DEBUG_TRAP_RECOUNT
        if (!ok) __trap();
#endif
        if (!ok) {
            // a 16-bit half wrapped: recount this CTA's part of the file exactly, straight into its (zeroed) row
            __threadfence();
            Gmem64Sink gs;
            gs.g = g;
            const uint64_t F0 = file_off[file], F1 = F0 + file_len[file];
            const GlobalSrc src{arena};
            const uint64_t lo_b = file_lo > F0 ? file_lo : F0, hi_b = file_hi < F1 ? file_hi : F1;
            const uint64_t span = (hi_b - lo_b + NWARPS - 1) / NWARPS;
            const uint64_t a0 = lo_b + (uint64_t)warp * span, a1 = (a0 + span < hi_b) ? a0 + span : hi_b;
            if (virt) {
                // byte ownership (k-mers whose first base lies in the piece): whole chunks per warp, the state at a warp's
                // first chunk from the exact backward scan of the range processor
                const uint64_t c_lo = lo_b / CHUNK, c_hi = (hi_b + CHUNK - 1) / CHUNK;
                const uint64_t per = (c_hi - c_lo + NWARPS - 1) / NWARPS;
                const uint64_t wc0 = c_lo + (uint64_t)warp * per, wc1 = wc0 + per < c_hi ? wc0 + per : c_hi;
                if (wc0 < wc1) {
                    const uint64_t o1 = wc1 * CHUNK < hi_b ? wc1 * CHUNK : hi_b;
                    fasta_process_range<7, false, 2>(src, (uint32_t)wc0, (uint32_t)wc1, (uint32_t)(F0 / CHUNK), gs, wc0 * CHUNK, o1, false);
                }
            } else if (a0 < hi_b) {
                const uint64_t lo = fasta_line_start_at_or_after(src, a0, F0, F1, lane);
                const uint64_t hi = (a1 >= hi_b && hi_b >= F1) ? F1 : fasta_line_start_at_or_after(src, a1, F0, F1, lane);
                lg_generic_region<7>(src, lo, hi, (uint32_t)(F0 / CHUNK), gs);
            }
        }
    };

    // the tile plan is cut for the generic kernel's grid; this CTA takes cta_stride consecutive shares of it.
    // Consecutive tiles of one file are contiguous in the arena: they are processed as ONE piece.
    const int t1 = cta_begin[(blockIdx.x + 1) * cta_stride];
    KF_T(tk0);
    for (int t = cta_begin[blockIdx.x * cta_stride]; t < t1;) {
        const Tile T = tiles[t];
        const uint32_t cur_P = file_P[T.file];
        if (ALL ? cur_P == 0u : cur_P != MYP) { ++t; continue; }
        uint32_t n_chunks = T.n_chunks;
        int te = t + 1;
        while (te < t1 && tiles[te].file == T.file && tiles[te].first_chunk == T.first_chunk + n_chunks) {
            n_chunks += tiles[te].n_chunks;
            ++te;
        }
        const uint64_t X0 = (uint64_t)T.first_chunk * CHUNK, X1 = X0 + (uint64_t)n_chunks * CHUNK;
        if ((int)T.file != cur_file) {
            if (cur_file >= 0) flush(cur_file);
            cur_file = (int)T.file;
            file_lo = X0;
        } else {
            __syncthreads();   // every warp is done with the previous piece (the cursor is about to be reset)
        }
        file_hi = X1;
        const uint64_t F0 = file_off[T.file], F1 = F0 + file_len[T.file];
        const uint64_t Xe = X1 < F1 ? X1 : F1;
        KF_T(ta0);
        if (threadIdx.x == 0) { *s_cursor = 0; *s_found = 0ull; }
        SingleSink gs;
        gs.h = single16;
        gs.n = s_nsingle;
        KF_T(ta1);
#ifdef KF_VL_TIMING
        const long long t_p0 = clock64();
#endif
        if (VIRT || (ALL && cur_P == KF_P_VIRTUAL)) {
            // the state at the piece's first byte.  A piece that begins inside the file: the last '\n' before X0 decides --
            // exact backward scan by the whole CTA, four chunks per warp and round (the line may be megabytes long)
            uint32_t st = 2u;
            if (X0 > F0) {
                // Groups of BS chunks counted backwards from X0; warp w takes the groups w, w + NWARPS, ...  No barrier in the
                // loop: a warp stops when it has found a '\n' (its later groups lie further back), when it has passed the
                // file's start, or when another warp has found one nearer to X0 than its next group.  Two groups in flight.
                const uint64_t c_first = F0 / CHUNK, c_top = X0 / CHUNK;   // chunks [c_first, c_top) lie before the piece
                constexpr int BS = 8;
                const uint64_t n_groups = (c_top - c_first + BS - 1) / BS;
                auto load_group = [&](uint64_t gi, uint4 (&w)[BS]) {
#pragma unroll
                    for (int u = 0; u < BS; u++) {
                        const uint64_t back = gi * BS + (uint64_t)u + 1;
                        w[u] = (gi < n_groups && c_top - c_first >= back) ? __ldg(reinterpret_cast<const uint4 *>(arena) + (c_top - back) * 32 + lane)
                                                                          : make_uint4(0u, 0u, 0u, 0u);
                    }
                };
                auto scan_group = [&](uint64_t gi, const uint4 (&w)[BS]) -> bool {   // true: found (and published)
                    uint32_t best = 0;   // 1 + offset (from F0) of the last '\n' this lane has seen, 0: none
#pragma unroll
                    for (int u = BS - 1; u >= 0; u--) {   // (u = 0 is the chunk nearest to X0: looked at last, so it wins)
                        const uint32_t m = newline_mask16(w[u]);
                        if (m) best = (uint32_t)((c_top - (gi * BS + (uint64_t)u + 1)) * CHUNK + (uint64_t)lane * 16 + (uint64_t)(31 - __clz((int)m)) - F0) + 1u;
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) { const uint32_t t = __shfl_xor_sync(FULL, best, o); best = t > best ? t : best; }
                    if (best && lane == 0) atomicMax(s_found, F0 + (uint64_t)best);   // = position of that '\n' + 1: a line start
                    return best != 0;
                };
                auto worth = [&](uint64_t gi) -> bool {   // may group gi still hold the nearest '\n'?
                    if (gi >= n_groups) return false;
                    unsigned long long fnd = 0;
                    if (lane == 0) fnd = *reinterpret_cast<volatile unsigned long long *>(s_found);
                    fnd = __shfl_sync(FULL, fnd, 0);                    // (one answer per warp)
                    return fnd < (c_top - gi * BS) * (uint64_t)CHUNK;   // (the group's bytes lie below that position)
                };
                __syncthreads();   // s_found = 0 is visible
                {
                    uint4 wa[BS], wb[BS];
                    uint64_t gi = (uint64_t)warp;
                    load_group(gi, wa);
                    for (;;) {
                        if (!worth(gi)) break;
                        load_group(gi + NWARPS, wb);
                        if (scan_group(gi, wa)) break;
                        gi += NWARPS;
                        if (!worth(gi)) break;
                        load_group(gi + NWARPS, wa);
                        if (scan_group(gi, wb)) break;
                        gi += NWARPS;
                    }
                }
                __syncthreads();
                const uint64_t ls = *s_found;   // start of the line that holds X0 (0: the file's first line, a header)
                st = ls == 0ull ? 1u : (ls == X0 ? 2u : (arena[ls] == (uint8_t)'>' ? 1u : 0u));
            }
            __syncthreads();
#ifdef KF_VL_TIMING
            const long long t_p1 = clock64();
#endif
            vl_process_piece<BASE>(arena, X0, Xe, st, F0, F1, buf, s_wscr + LN_SCR_WORDS * warp, bar, par, s_cursor, (uint32_t)NWARPS, hist16, gs, npairs, vlog);
#ifdef KF_VL_TIMING
            __syncthreads();
            if (threadIdx.x == 0 && blockIdx.x < 160) {
                const long long t_p2 = clock64();
                g_vl_cta[4 * blockIdx.x + 0] += (unsigned long long)(t_p2 - t_p0);
                g_vl_cta[4 * blockIdx.x + 1] += 1;
                if ((unsigned long long)(t_p2 - t_p1) > g_vl_cta[4 * blockIdx.x + 2]) g_vl_cta[4 * blockIdx.x + 2] = (unsigned long long)(t_p2 - t_p1);
                g_vl_cta[4 * blockIdx.x + 3] += (unsigned long long)(t_p1 - t_p0);
            }
#endif
        } else {
        // every warp finds the piece's anchor (its first line start) by itself: same loads, served by L1 after the first
        const uint64_t A = fasta_line_start_at_or_after(GlobalSrc{arena}, X0, F0, F1, lane);
        __syncthreads();
#define KF_LN_PIECE(W) ln_process_piece<W, BASE>(arena, A, Xe, F0, F1, buf, s_wscr + LN_SCR_WORDS * warp, bar, par, s_cursor, (uint32_t)NWARPS, hist16, gs, npairs)
        if constexpr (ALL) {
            switch (cur_P) {   // (uniform over the CTA)
                case 81: KF_LN_PIECE(80); break;
                case 61: KF_LN_PIECE(60); break;
                case 71: KF_LN_PIECE(70); break;
                case 101: KF_LN_PIECE(100); break;
                case 51: KF_LN_PIECE(50); break;
                default: break;
            }
        } else {
            KF_LN_PIECE(LW);
        }
#undef KF_LN_PIECE
        }
        KF_T(ta2);
        KF_TADD(0, ta2 - ta1);
        t = te;
    }
    if (cur_file >= 0) flush(cur_file);
#ifdef KF_PIECE_TIMING
    if (threadIdx.x == 0) atomicAdd(&g_piece_timing[4], (unsigned long long)(clock64() - tk0));
#endif
}

#ifndef KF_EMU
// Shared-window address at which the dynamic shared memory of a kernel WITHOUT static shared memory begins, under the
// launch conditions of this process (0x400 on sm_100: the first KiB is the system's; debuggers / sanitizers may differ).
__global__ void kf_smem_base_probe_kernel(uint32_t *out) {
    KF_DYN_SMEM(uint32_t, probe_smem);
    if (threadIdx.x == 0) *out = smem_addr(probe_smem);
}
#endif

// Line width of each FASTA file, judged from its first lines: P = LW + 1 if the three lines after the
// header are LW bases wide (LW one of 60/70/80), else 0 (generic kernel).  One warp per file: 512-byte chunks,
// newline masks per lane, the first four '\n' positions picked out of the ballots.
__global__ void __launch_bounds__(128)
probe_line_width_kernel(const uint8_t *__restrict__ arena, const uint64_t *__restrict__ file_off,
                        const uint64_t *__restrict__ file_len, const uint8_t *__restrict__ formats, int n,
                        uint32_t force_generic, uint32_t *__restrict__ file_P,
                        uint32_t *__restrict__ width_counts /* [line_width_slot()]: 0 generic, 1..3 60/70/80, 4 long lines, 5 100, 7 50 */,
                        unsigned long long *__restrict__ g_fwd, const uint32_t *__restrict__ file_row, uint32_t row_bins) {
    const int lane = threadIdx.x & 31;
    const int f = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (f >= n) return;
    uint32_t P = 0;
    if (!force_generic && formats[f] == (uint8_t)'>') {
        const uint64_t F0 = file_off[f], L = file_len[f];
        const uint64_t nchunks = (L + CHUNK - 1) / CHUNK < 16 ? (L + CHUNK - 1) / CHUNK : 16;   // look at most 8 KiB in
        uint64_t nlpos[4];
        int found = 0;
        for (uint64_t c = 0; c < nchunks && found < 4; c++) {
            const uint4 w = __ldg(reinterpret_cast<const uint4 *>(arena + F0) + c * 32 + lane);
            const uint64_t pb = c * CHUNK + (uint64_t)lane * 16;   // offset in the file
            uint32_t m = newline_mask16(w);
            if (pb + 16 > L) m &= pb >= L ? 0u : ((1u << (uint32_t)(L - pb)) - 1u);
            unsigned Bm = __ballot_sync(FULL, m != 0);
            while (Bm && found < 4) {
                const int j = __ffs((int)Bm) - 1;
                Bm &= Bm - 1;
                uint32_t mj = __shfl_sync(FULL, m, j);
                while (mj && found < 4) {
                    nlpos[found++] = c * CHUNK + (uint64_t)j * 16 + (uint32_t)(__ffs((int)mj) - 1);
                    mj &= mj - 1;
                }
            }
        }
        if (found == 4) {
            const uint64_t w0 = nlpos[1] - nlpos[0] - 1, w1 = nlpos[2] - nlpos[1] - 1, w2 = nlpos[3] - nlpos[2] - 1;
            if (w0 == w1 && w1 == w2 && line_width_slot((int)w0) != 0) P = (uint32_t)w0 + 1;
        }
        // long lines (one line per contig, as assemblers write them): fewer than four '\n' in the first 8 KiB, or a line of
        // 2 KiB and more among the first ones -> virtual lines (exact for any structure; fast when '\n' are rare)
        if (P == 0 && L >= 32768) {
            bool longl = found < 4;
            for (int i = 1; i < found; i++) longl = longl || nlpos[i] - nlpos[i - 1] > 2048;
            if (longl) P = KF_P_VIRTUAL;
        }
    }
    if (lane == 0) {
        file_P[f] = P;
        atomicAdd(width_counts + (P == KF_P_VIRTUAL ? 4 : P ? line_width_slot((int)P - 1) : 0), 1u);
    }
    // Forward rows: the line kernel WRITES every bin of the rows of the files it takes; every other file's rows are
    // added to with atomics (generic / FASTQ kernels) or never touched (unsupported input), so they are zeroed here --
    // this replaces a memset of the whole workspace.
    if (P == 0 && g_fwd != nullptr) {
        ulonglong2 *row = reinterpret_cast<ulonglong2 *>(g_fwd + (size_t)file_row[f] * row_bins);
        const size_t n16 = (size_t)(file_row[f + 1] - file_row[f]) * row_bins / 2;
        for (size_t i = lane; i < n16; i += 32) row[i] = make_ulonglong2(0ull, 0ull);
    }
}

// ================================================================================================
// FASTQ (4-line records): every lane chases its own records
// ================================================================================================
// Header and quality lines may hold ANY byte but '\n' (qualities legitimately contain A/C/G/T and may begin with '@'
// or '+'), so a record's structure cannot be read off isolated bytes.  What 4-line FASTQ does guarantee -- and what
// Jellyfish relies on as well -- is that the quality line is exactly as long as the sequence line.  So a lane that
// stands on a record start can walk the file alone: header line up to its '\n', sequence line (counted), '+' line up to
// its '\n', then JUMP over the quality line by length.  More than half of the file (qualities) is never even loaded.
//   work split : a tile = 32 lane ranges of FQ_LANE_BYTES; a lane owns the records whose header starts in its range and
//                finishes the last one beyond the range end.
//   sync       : a range start is put on a record boundary with a rule that is exact for 4-line FASTQ: the first line
//                start s with byte '@' whose second-next line starts with '+'.  (A quality line that begins with '@' is
//                followed by a header and then a SEQUENCE line, which never begins with '+'.)
//   inner loop : one 16-byte piece per iteration, the same code for every lane whatever its state (header / sequence /
//                plus line), so the lanes of a warp stay converged although they sit in different records: decode 16
//                bytes + 8 look-ahead bytes, byte masks for "not A/C/G/T" and '\n', a few state transitions, then the
//                k-mer that starts at byte j counts iff bad[j .. j+K) is clear.
//   layout     : the line after a sequence line must begin with '+', the byte after the skipped quality line must be
//                '\n' followed by '@' (or the end of the file); the lowest offending offset per file lands in fq_err and
//                the host maps it to KF_ERR_FASTQ.
constexpr uint32_t FQ_LANE_BYTES = 2048;

// position of the first '\n' at or after p (p < end), or end.  16-byte pieces of the arena.
__device__ __forceinline__ uint64_t fq_next_newline(const uint8_t *__restrict__ arena, uint64_t p, uint64_t end) {
    while (p < end) {
        const uint64_t a = p & ~15ull;
        const uint4 w = __ldg(reinterpret_cast<const uint4 *>(arena + a));
        uint32_t m = newline_mask16(w) & ~((1u << (uint32_t)(p - a)) - 1u);
        if (m) {
            const uint64_t q = a + (uint32_t)(__ffs((int)m) - 1);
            return q < end ? q : end;
        }
        p = a + 16;
    }
    return end;
}

// Same, for a line that starts at p: also checks the line's first byte (`first`: '@' or '+') and, when `prev_nl` is set,
// that the byte before p is a '\n' (the end of the quality line that was skipped by length).  Returns ~0 on a violation.
__device__ __forceinline__ uint64_t fq_line_end_checked(const uint8_t *__restrict__ arena, uint64_t p, uint64_t end, uint32_t first,
                                                        bool prev_nl) {
    const uint64_t a0 = p & ~15ull;
    const uint32_t off = (uint32_t)(p - a0);
    uint4 w = __ldg(reinterpret_cast<const uint4 *>(arena + a0));
    if (byte_of(w, (int)off) != first) return ~0ull;
    if (prev_nl) {
        const uint32_t pb = off ? byte_of(w, (int)off - 1) : (uint32_t)arena[p - 1];
        if (pb != 0x0Au) return ~0ull;
    }
    uint32_t m = newline_mask16(w) & ~((1u << off) - 1u);
    uint64_t a = a0;
    for (;;) {
        if (m) {
            const uint64_t q = a + (uint32_t)(__ffs((int)m) - 1);
            return q < end ? q : end;
        }
        a += 16;
        if (a >= end) return end;
        w = __ldg(reinterpret_cast<const uint4 *>(arena + a));
        m = newline_mask16(w);
    }
}

// A sink may take a FASTQ piece's 16 positions at once (fq16(hi, lo, ok): bit j of ok = the k-mer at base j counts).
template <class S> struct sink_takes_fq16 { static constexpr bool value = false; };

// k = 7 FASTQ: 7-mers counted as 8-mers ("pairs", as the line kernel does): the 8-mer at an even base j of the piece is the
// 7-mers j and j + 1 -- half the shared-memory atomics, which is what a FASTQ warp's time is made of.  Pair histogram:
// 65,536 8-mer bins (first base MOST significant here) as 32,768 words of two u16 halves, word = v >> 1, low half = every
// pair of the word, high half = the odd ones; a 7-mer whose partner is not valid (read ends, N) goes to a u32 singles
// histogram in a rarely taken branch.  npairs / the low-half checksum catch a wrapped half (flush: exact recount).
struct FqPairSink {
    uint32_t *hist16, *single;
    uint32_t hbase;          // shared-window address of hist16 (0 in the emulation)
    uint32_t npairs;         // pairs issued by this thread since the last flush
    __device__ __forceinline__ void fq16(uint32_t hi, uint32_t lo, uint32_t ok) {
        const uint32_t pm = ok & (ok >> 1) & 0x5555u;          // even j: 7-mers j and j + 1 both count
        npairs += (uint32_t)__popc(pm);
        if (pm) {
            // no branch per pair: a pair that does not count adds 0 (its address is a valid bin all the same)
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
                const uint32_t off8 = kmer_off_at<8>(hi, lo, j);   // 4 * v
                const uint32_t vm = 0u - ((pm >> j) & 1u);
                red_shared_add(hist16, hbase, (off8 >> 1) & 0x1FFFCu, (((off8 & 4u) << 14) | 1u) & vm);
            }
        }
        uint32_t sm = ok & ~(pm | (pm << 1));                  // counted 7-mers outside a pair
        while (sm) {
            const int j = __ffs((int)sm) - 1;
            sm &= sm - 1;
            const unsigned long long w = ((unsigned long long)hi << 32) | lo;
            atomicAdd(single + ((uint32_t)(w >> (50 - 2 * j)) & 0x3FFFu), 1u);
        }
    }
    __device__ __forceinline__ void operator()(uint32_t) const {}
};
template <> struct sink_takes_fq16<FqPairSink> { static constexpr bool value = true; };
struct FqRecountSink {   // exact recount after a wrapped half: the two 7-mers of every pair, one global atomic each
    unsigned long long *g;   // the file's u64 row
    __device__ __forceinline__ void fq16(uint32_t hi, uint32_t lo, uint32_t ok) const {
        uint32_t pm = ok & (ok >> 1) & 0x5555u;
        pm |= pm << 1;
        const unsigned long long w = ((unsigned long long)hi << 32) | lo;
        while (pm) {
            const int j = __ffs((int)pm) - 1;
            pm &= pm - 1;
            atomicAdd(g + ((uint32_t)(w >> (50 - 2 * j)) & 0x3FFFu), 1ull);
        }
    }
    __device__ __forceinline__ void operator()(uint32_t) const {}
};
template <> struct sink_takes_fq16<FqRecountSink> { static constexpr bool value = true; };

// k-mers that start in one 16-byte piece: hi = its bases, lo = the next piece's, bad32 = bad bits of both (piece in the
// low half); the k-mer at byte j counts iff bad[j .. j+K) is clear.
template <int K, class Sink>
__device__ __forceinline__ void fq_emit16(uint32_t hi, uint32_t lo, uint32_t bad32, Sink &sink) {
    uint32_t o = bad32;
    int cover = 1;
#pragma unroll
    for (int it = 0; it < 4; it++) {
        if (cover < K) {
            const int sft = (cover < K - cover) ? cover : K - cover;
            o |= o >> sft;
            cover += sft;
        }
    }
    const uint32_t ok = ~o & 0xFFFFu;
    if constexpr (sink_takes_fq16<Sink>::value) {
        sink.fq16(hi, lo, ok);
    } else if (ok) {
#pragma unroll
        for (int j = 0; j < 16; j++) {
            if constexpr (sink_has_add_if<Sink>::value) sink.add_if(kmer_off_at<K>(hi, lo, j), ok, 1u << j);
            else if ((ok >> j) & 1u) sink(kmer_off_at<K>(hi, lo, j));
        }
    }
}

// One lane: count the k-mers of every record whose header line starts in [X0, X1) of the file [F0, F1).
// The lanes of a warp walk their records in step: header lines together, sequence lines together (16-byte pieces, the
// previous piece is counted when the next one -- its look-ahead -- has been decoded), plus lines together, jump.
template <int K, class Sink>
__device__ __forceinline__ void fastq_lane_records(const uint8_t *__restrict__ arena, uint64_t F0, uint64_t F1, uint64_t X0,
                                                   uint64_t X1, Sink &sink, unsigned long long *fq_err) {
    if (X0 >= F1) X0 = X1 = F1;
    if (X1 > F1) X1 = F1;
    // ---- put the range start on a record boundary ----
    uint64_t pos = X0;          // a record (header line) start, once synchronised
    bool run = X0 < X1;
    if (run && X0 > F0) {
        // line starts after X0 - 1: s0, s1, s2 ...; the first s_i that holds '@' while s_(i+2) holds '+'
        uint64_t s0 = fq_next_newline(arena, X0 - 1, F1) + 1;
        uint64_t s1 = s0 < F1 ? fq_next_newline(arena, s0, F1) + 1 : F1 + 1;
        run = false;
        for (int it = 0; it < 8; it++) {
            if (s0 >= X1 || s0 >= F1) break;
            const uint64_t s2 = s1 < F1 ? fq_next_newline(arena, s1, F1) + 1 : F1 + 1;
            // near the end of the file the second-next line may not exist: then s0 is a header iff the line after it
            // (if any) is a sequence line, i.e. does not begin with '@'
            const bool hdr = arena[s0] == (uint8_t)'@' &&
                             (s2 < F1 ? arena[s2] == (uint8_t)'+' : (s1 >= F1 || arena[s1] != (uint8_t)'@'));
            if (hdr) { pos = s0; run = true; break; }
            s0 = s1;
            s1 = s2;
        }
    }
    bool after_jump = false;   // pos was reached by skipping a quality line: the byte before it must be its '\n'
    while (run && pos < F1) {
        // ---- header line ----
        const uint64_t hn = fq_line_end_checked(arena, pos, F1, (uint32_t)'@', after_jump);
        if (hn == ~0ull) { atomicMin(fq_err, pos); break; }
        if (hn >= F1) break;
        const uint64_t s = hn + 1;   // first byte of the sequence line
        // ---- sequence line ----
        uint64_t a = s & ~15ull;
        uint4 w = __ldg(reinterpret_cast<const uint4 *>(arena + a));
        uint32_t lead = (1u << (uint32_t)(s - a)) - 1u;   // bytes of the first piece that lie before the sequence
        uint32_t pbits = 0, pbad = 0xFFFFu, plus_byte = 0;
        bool have_prev = false, plus_known = false;
        uint64_t seq_end;
        for (;;) {
            const uint4 wn = __ldg(reinterpret_cast<const uint4 *>(arena + a + 16));   // next piece, asked for early
            uint32_t p0, p1, p2, p3, V0, V1, V2, V3;
            decode_word(w.x, p0, V0);
            decode_word(w.y, p1, V1);
            decode_word(w.z, p2, V2);
            decode_word(w.w, p3, V3);
            const uint32_t r1 = __byte_perm(p3, p2, 0x0073);
            const uint32_t r2 = __byte_perm(p1, p0, 0x0073);
            const uint32_t bits = __byte_perm(r1, r2, 0x5410);
            uint32_t inv = 0, nlm = 0;
            if (V0 | V1 | V2 | V3) {
                inv = movemask4(nonzero_bytes(V0)) | (movemask4(nonzero_bytes(V1)) << 4) | (movemask4(nonzero_bytes(V2)) << 8) |
                      (movemask4(nonzero_bytes(V3)) << 12);
                // a '\n' has bit 6 clear; a piece of letters only (bases and N: the usual reason to be here) holds none
                if (~(w.x & w.y & w.z & w.w) & 0x40404040u) nlm = newline_mask16(w);
            }
            if (a + 16 > F1) { const uint32_t v = (uint32_t)(F1 - a); inv |= 0xFFFFu & ~((1u << v) - 1u); nlm |= 1u << v; }   // file end = line end
            nlm &= ~lead;
            const uint32_t e = nlm ? (uint32_t)(__ffs((int)nlm) - 1) : 16u;   // the line's end inside this piece
            const uint32_t bad = (inv | lead | (0xFFFFu & ~((1u << e) - 1u))) & 0xFFFFu;
            if (have_prev) fq_emit16<K>(pbits, bits, pbad | (bad << 16), sink);
            have_prev = true;
            pbits = bits;
            pbad = bad;
            lead = 0;
            if (e < 16u) {
                seq_end = a + e;
                if (e < 15u) { plus_byte = byte_of(w, (int)e + 1); plus_known = true; }
                break;
            }
            a += 16;
            w = wn;
        }
        fq_emit16<K>(pbits, 0u, pbad | 0xFFFF0000u, sink);
        const uint32_t seqlen = (uint32_t)(seq_end - s);
        KF_PREFETCH_L2(arena + seq_end + 4 + seqlen);   // where the next record starts if the plus line is bare ("+\n")
        // ---- plus line, then jump over the quality line ----
        const uint64_t pp = seq_end + 1;
        if (pp >= F1) break;
        if (plus_known && plus_byte != (uint32_t)'+') { atomicMin(fq_err, pp); break; }
        const uint64_t pn = fq_line_end_checked(arena, pp, F1, (uint32_t)'+', false);
        if (pn == ~0ull) { atomicMin(fq_err, pp); break; }
        if (pn >= F1) break;
        pos = pn + 1 + seqlen + 1;   // quality line: seqlen bytes and its '\n'
        after_jump = true;
        if (pos >= X1) {
            // that record belongs to the next range, whose lane finds its start by itself: the jump is validated here
            // (the skipped quality line must end right before pos, and a header must start there)
            if (pos < F1 && (arena[pos - 1] != 0x0Au || arena[pos] != (uint8_t)'@')) atomicMin(fq_err, pos);
            else if (pos == F1 && arena[pos - 1] != 0x0Au) atomicMin(fq_err, pos - 1);
            break;
        }
    }
}

// FASTQ counting, k <= 7: per-CTA shared-memory histogram; CTA b owns tiles [cta_begin[b], cta_begin[b+1]), its warps take
// them from a shared counter, the histogram is flushed when the CTA moves to another file.
template <int K, int THREADS, int MIN_CTAS>
__global__ void __launch_bounds__(THREADS, MIN_CTAS)
count_fastq_smem_kernel(const uint8_t *__restrict__ arena, const Tile *__restrict__ tiles, const int *__restrict__ cta_begin,
                        const uint64_t *__restrict__ file_off, const uint64_t *__restrict__ file_len,
                        unsigned long long *__restrict__ g_fwd, const uint32_t *__restrict__ file_row,
                        unsigned long long *__restrict__ fq_err) {
    KF_DYN_SMEM(uint32_t, hist);
    constexpr int NB = 1 << (2 * K);
    __shared__ uint32_t s_next;
    for (int i = threadIdx.x; i < NB; i += THREADS) hist[i] = 0;
    if (threadIdx.x == 0) s_next = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const SmemSink emit = make_smem_sink(hist);
    const int t1 = cta_begin[blockIdx.x + 1];
    for (int t = cta_begin[blockIdx.x]; t < t1;) {
        const uint32_t file = tiles[t].file;
        int te = t + 1;
        while (te < t1 && tiles[te].file == file) ++te;
        const uint64_t F0 = file_off[file], F1 = F0 + file_len[file];
        for (;;) {
            int tt = 0;
            if (lane == 0) tt = t + (int)atomicAdd(&s_next, 1u);
            tt = __shfl_sync(FULL, tt, 0);
            if (tt >= te) break;
            const Tile T = tiles[tt];
            const uint64_t tb = (uint64_t)T.first_chunk * CHUNK, tend = tb + (uint64_t)T.n_chunks * CHUNK;
            const uint64_t x0 = tb + (uint64_t)lane * FQ_LANE_BYTES;
            const uint64_t x1 = x0 + FQ_LANE_BYTES < tend ? x0 + FQ_LANE_BYTES : tend;
            if (x0 < tend) fastq_lane_records<K>(arena, F0, F1, x0, x1, emit, fq_err + file);
        }
        __syncthreads();
        if (threadIdx.x == 0) s_next = 0;
        unsigned long long *g = g_fwd + (size_t)file_row[file] * NB;
        for (int i = threadIdx.x; i < NB; i += THREADS) {
            const uint32_t v = hist[i];
            if (v) { atomicAdd(g + i, (unsigned long long)v); hist[i] = 0; }
        }
        __syncthreads();
        t = te;
    }
}

// FASTQ counting, k = 7: the pair histogram (FqPairSink).  One CTA per SM (192 KB of histograms); CTA b owns the tile ranges
// [cta_begin[b * stride], cta_begin[(b + 1) * stride]) of the plan, its warps take tiles from a shared counter; flush when
// the CTA moves to another file: checksum, 7-mer counts out of the pair + singles histograms, u64 atomics into the file's
// row (several CTAs may hold tiles of one file).
template <int THREADS>
__global__ void __launch_bounds__(THREADS, 1)
count_fastq_pairs_kernel(const uint8_t *__restrict__ arena, const Tile *__restrict__ tiles, const int *__restrict__ cta_begin, int stride,
                         const uint64_t *__restrict__ file_off, const uint64_t *__restrict__ file_len,
                         unsigned long long *__restrict__ g_fwd, const uint32_t *__restrict__ file_row,
                         unsigned long long *__restrict__ fq_err) {
    constexpr int K = 7, NB = 16384, NWORDS = 32768, NWARPS = THREADS / 32;
    KF_DYN_SMEM(uint32_t, smem);
    uint32_t *hist16 = smem, *single = smem + NWORDS;
    uint32_t *s_part = single + NB;         // [2 * NWARPS] pairs issued | low-half sums
    uint32_t *s_next = s_part + 2 * NWARPS;
    for (int i = threadIdx.x; i < NWORDS + NB; i += THREADS) smem[i] = 0;
    if (threadIdx.x == 0) *s_next = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    FqPairSink sink;
    sink.hist16 = hist16;
    sink.single = single;
    sink.hbase = smem_addr(hist16);
    sink.npairs = 0;
    const int t1 = cta_begin[(blockIdx.x + 1) * stride];
    for (int t = cta_begin[blockIdx.x * stride]; t < t1;) {
        const uint32_t file = tiles[t].file;
        int te = t + 1;
        while (te < t1 && tiles[te].file == file) ++te;
        const uint64_t F0 = file_off[file], F1 = F0 + file_len[file];
        for (;;) {
            int tt = 0;
            if (lane == 0) tt = t + (int)atomicAdd(s_next, 1u);
            tt = __shfl_sync(FULL, tt, 0);
            if (tt >= te) break;
            const Tile T = tiles[tt];
            const uint64_t tb = (uint64_t)T.first_chunk * CHUNK, tend = tb + (uint64_t)T.n_chunks * CHUNK;
            const uint64_t x0 = tb + (uint64_t)lane * FQ_LANE_BYTES;
            const uint64_t x1 = x0 + FQ_LANE_BYTES < tend ? x0 + FQ_LANE_BYTES : tend;
            if (x0 < tend) fastq_lane_records<K>(arena, F0, F1, x0, x1, sink, fq_err + file);
        }
        // ---- flush ----
        uint32_t np = sink.npairs;
        sink.npairs = 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) np += __shfl_xor_sync(FULL, np, o);
        if (lane == 0) s_part[warp] = np;
        __syncthreads();
        if (threadIdx.x == 0) *s_next = 0;
        uint32_t low = 0;
        for (int i = threadIdx.x; i < NWORDS; i += THREADS) low += hist16[i] & 0xFFFFu;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) low += __shfl_xor_sync(FULL, low, o);
        if (lane == 0) s_part[NWARPS + warp] = low;
        __syncthreads();
        unsigned long long tp = 0, tl = 0;
        for (int w = 0; w < NWARPS; w++) { tp += s_part[w]; tl += s_part[NWARPS + w]; }
        const bool ok = tp == tl;
        unsigned long long *g = g_fwd + (size_t)file_row[file] * NB;
        // 7-mer x (first base most significant): first of the 8-mers 4x + b (words 2x, 2x + 1: every pair of both), second
        // of the 8-mers a * 16384 + x (word a * 8192 + (x >> 1): the odd half for odd x, low - high for even x)
        for (int j = threadIdx.x; j < NB / 2; j += THREADS) {
            uint32_t c0 = single[2 * j], c1 = single[2 * j + 1];
            if (ok) {
                const uint4 pw = *reinterpret_cast<const uint4 *>(hist16 + 4 * j);
                c0 += (pw.x & 0xFFFFu) + (pw.y & 0xFFFFu);
                c1 += (pw.z & 0xFFFFu) + (pw.w & 0xFFFFu);
#pragma unroll
                for (int a = 0; a < 4; a++) {
                    const uint32_t w = hist16[a * 8192 + j];
                    c0 += (w & 0xFFFFu) - (w >> 16);
                    c1 += w >> 16;
                }
            }
            if (c0) atomicAdd(g + 2 * j, (unsigned long long)c0);
            if (c1) atomicAdd(g + 2 * j + 1, (unsigned long long)c1);
        }
        __syncthreads();
        for (int i = threadIdx.x; i < NWORDS + NB; i += THREADS) smem[i] = 0;
        __syncthreads();
        if (!ok) {
            // a 16-bit half wrapped: this CTA's tiles of the file again, the pairs' 7-mers straight into the row.  (The
            // singles were added above; FqRecountSink counts only what the pair histogram held.)
            FqRecountSink rs;
            rs.g = g;
            for (int tt = t + warp; tt < te; tt += NWARPS) {
                const Tile T = tiles[tt];
                const uint64_t tb = (uint64_t)T.first_chunk * CHUNK, tend = tb + (uint64_t)T.n_chunks * CHUNK;
                const uint64_t x0 = tb + (uint64_t)lane * FQ_LANE_BYTES;
                const uint64_t x1 = x0 + FQ_LANE_BYTES < tend ? x0 + FQ_LANE_BYTES : tend;
                if (x0 < tend) fastq_lane_records<K>(arena, F0, F1, x0, x1, rs, fq_err + file);
            }
            __syncthreads();
        }
        t = te;
    }
}

// FASTQ counting, k >= 8: forward counts in global memory.
template <int K, int THREADS>
__global__ void __launch_bounds__(THREADS)
count_fastq_gmem_kernel(const uint8_t *__restrict__ arena, const Tile *__restrict__ tiles, const int *__restrict__ cta_begin,
                        const uint64_t *__restrict__ file_off, const uint64_t *__restrict__ file_len,
                        uint32_t *__restrict__ g_fwd32, uint32_t file_base, unsigned long long *__restrict__ fq_err) {
    constexpr size_t NB = (size_t)1 << (2 * K);
    constexpr int NWARPS = THREADS / 32;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t1 = cta_begin[blockIdx.x + 1];
    for (int t = cta_begin[blockIdx.x] + warp; t < t1; t += NWARPS) {
        const Tile T = tiles[t];
        GmemSink emit;
        emit.g = g_fwd32 + (size_t)(T.file - file_base) * NB;
        const uint64_t F0 = file_off[T.file], F1 = F0 + file_len[T.file];
        const uint64_t tb = (uint64_t)T.first_chunk * CHUNK, tend = tb + (uint64_t)T.n_chunks * CHUNK;
        const uint64_t x0 = tb + (uint64_t)lane * FQ_LANE_BYTES;
        const uint64_t x1 = x0 + FQ_LANE_BYTES < tend ? x0 + FQ_LANE_BYTES : tend;
        if (x0 < tend) fastq_lane_records<K>(arena, F0, F1, x0, x1, emit, fq_err + T.file);
    }
}

// ------------------------------------------------------------------------------------------------
// FASTQ that is not in 4-line layout (sequences and qualities wrapped over several lines)
// ------------------------------------------------------------------------------------------------
// Record boundaries of multi-line FASTQ cannot be found locally ('@' and '+' are legal first quality characters), so such
// a file is read the way Jellyfish reads it -- front to back: header line, sequence lines up to the line that starts
// with '+', then quality lines until as many quality characters as bases have gone by (oracle/kf_oracle.c walk_fastq).
// One warp per file, exact and slow (a file the record-chasing kernel above reported in fq_err, and only that, comes
// here): the warp finds each line's end together, and counts the k-mers that END in a sequence line from the stream
// "last k-1 bytes of the record so far" + line, so k-mers run on over the line ends of a record.  Row: the file's forward
// counts (u64 for k <= 7, u32 above), zeroed first -- the fast kernel has left partial counts in it.
template <typename RowT>
__global__ void __launch_bounds__(128)
fastq_multiline_kernel(const uint8_t *__restrict__ arena, const uint64_t *__restrict__ file_off, const uint64_t *__restrict__ file_len,
                       const uint8_t *__restrict__ formats, uint32_t f0, uint32_t f1, int k, RowT *__restrict__ g_fwd,
                       const uint32_t *__restrict__ file_row, unsigned long long *__restrict__ fq_err) {
    __shared__ uint8_t s_tail[4][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t f = f0 + (uint32_t)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (f >= f1 || formats[f] != (uint8_t)'@') return;
    const uint64_t F0 = file_off[f], F1 = F0 + file_len[f];
    const unsigned long long e = fq_err[f];
    if (e == ~0ull || e < F0 || e >= F1) return;                       // 4-line layout held
    {
        // a violation followed by line ends only (trailing blank lines ...) is none
        bool other = false;
        for (uint64_t p = e + lane; p < F1 && p < e + 4096; p += 32) other = other || (arena[p] != 0x0Au && arena[p] != 0x0Du);
        if (F1 - e > 4096) other = true;
        if (__ballot_sync(FULL, other) == 0u) { if (lane == 0) fq_err[f] = ~0ull; return; }
    }
    const size_t NB = (size_t)1 << (2 * k);
    RowT *row = g_fwd + (file_row ? (size_t)file_row[f] : (size_t)(f - f0)) * NB;
    for (size_t i = lane; i < NB; i += 32) row[i] = (RowT)0;
    __threadfence();
    KF_SYNCWARP();
    const GlobalSrc gsrc{arena};
    uint8_t *tail = s_tail[warp];
    const int T = k - 1;                                               // bytes of the record's stream kept between lines
    auto line_end = [&](uint64_t q) -> uint64_t {                      // position of the first '\n' at or after q, F1 if none
        if (q >= F1) return F1;
        const uint64_t ls = fasta_line_start_at_or_after(gsrc, q + 1, F0, F1, lane);
        return (ls == F1 && arena[F1 - 1] != 0x0Au) ? F1 : ls - 1;
    };
    auto skip_nl = [&](uint64_t p) -> uint64_t { while (p < F1 && arena[p] == 0x0Au) p++; return p; };
    auto ignore_line = [&](uint64_t p) -> uint64_t { const uint64_t q = line_end(p); return q < F1 ? q + 1 : F1; };
    auto reset_tail = [&]() { KF_SYNCWARP(); if (lane < 32) tail[lane] = 0; KF_SYNCWARP(); };
    uint64_t p = ignore_line(F0);                                      // the first '@' header
    while (p < F1) {
        uint64_t nseq = 0, nq = 0;
        p = skip_nl(p);
        reset_tail();
        while (p < F1 && arena[p] != (uint8_t)'+') {
            const uint64_t q = line_end(p);
            const uint64_t len = q - p;
            // stream = tail[0 .. T) + line[0 .. len): the k-mers that start at stream positions [0, len) end in this line
            for (uint64_t s0 = lane; s0 < len; s0 += 32) {
                bool valid = true;
                uint32_t kmer = 0;
                for (int t = 0; t < k; t++) {
                    const uint64_t i = s0 + (uint64_t)t;
                    const uint32_t c = i < (uint64_t)T ? tail[i] : arena[p + (i - (uint64_t)T)];
                    valid = valid && is_base(c);
                    kmer = (kmer << 2) | ((c >> 1) & 3u);
                }
                if (valid) atomicAdd(row + kmer, (RowT)1);
            }
            KF_SYNCWARP();
            // the new tail: the last T bytes of the stream
            uint8_t nt = 0;
            if (lane < T) {
                const uint64_t i = len + (uint64_t)lane;               // stream position of the new tail's byte `lane`
                nt = i < (uint64_t)T ? tail[i] : arena[p + (i - (uint64_t)T)];
            }
            KF_SYNCWARP();
            if (lane < T) tail[lane] = nt;
            KF_SYNCWARP();
            nseq += len;
            p = skip_nl(q);
        }
        if (p >= F1) break;
        p = ignore_line(p);                                            // the '+' line
        p = skip_nl(p);
        while (p < F1 && nq < nseq) {
            const uint64_t q = line_end(p);
            nq += q - p;
            p = skip_nl(q);
        }
        p = skip_nl(p);
        p = ignore_line(p);                                            // the next '@' header
    }
    __threadfence();
    if (lane == 0) fq_err[f] = ~0ull;                                  // counted: no error to report
}

// ------------------------------------------------------------------------------------------------
// Chunked-genome mode (kf2vec/main.py:813-881): every sliding window of a linearised contig becomes a one-record
// pseudo-file ">\n<window bytes>" in a scratch arena, so the counting kernels above see it as an ordinary input
// and each window yields one row.  One CTA per window; src windows may overlap and need no alignment.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gather_windows_kernel(const uint8_t *__restrict__ seq, const uint64_t *__restrict__ win_off, const uint32_t *__restrict__ win_len,
                      int n, uint64_t slot_bytes, uint8_t *__restrict__ arena) {
    for (int w = blockIdx.x; w < n; w += gridDim.x) {
        uint8_t *dst = arena + (uint64_t)w * slot_bytes;
        const uint8_t *src = seq + win_off[w];
        const uint32_t len = win_len[w];
        if (threadIdx.x == 0) { dst[0] = (uint8_t)'>'; dst[1] = 0x0Au; }
        for (uint32_t i = threadIdx.x; i < len; i += blockDim.x) { const uint8_t c = src[i]; dst[2 + i] = c == 0x0Au ? (uint8_t)'N' : c; }   // a stray '\n' must break, not join
        for (uint64_t i = 2 + (uint64_t)len + threadIdx.x; i < slot_bytes; i += blockDim.x) dst[i] = 0;
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

// k <= 7: the file's rows are summed with coalesced loads into shared memory (4^k entries: u32 when no bin of the
// batch can reach 2^32, i.e. every file is shorter than 4 GiB -- 64 KB at k = 7, three CTAs per SM -- else u64), then
// the canonical gather reads from there.  A thread keeps its (at most FOLD_ITEMS) canonical counts in registers
// between the total and the normalisation.
#ifdef KF_EMU
constexpr int FOLD_THREADS = 64;   // (one OS thread per CUDA thread in the emulation)
constexpr int FOLD_ITEMS = 128;
#else
constexpr int FOLD_THREADS = 512;
constexpr int FOLD_ITEMS = 16;   // 8,192 canonical 7-mers / 512 threads
#endif
template <typename SmT>
__global__ void __launch_bounds__(FOLD_THREADS)
fold_normalize_smem_kernel(const unsigned long long *__restrict__ g_fwd, const uint32_t *__restrict__ canon, int k, long long V,
                           uint32_t flags, const uint32_t *__restrict__ file_P, const uint32_t *__restrict__ file_row,
                           unsigned long long *__restrict__ counts, double *__restrict__ freq, float *__restrict__ feat,
                           unsigned long long *__restrict__ totals) {
    KF_DYN_SMEM(unsigned long long, fold_smem);
    SmT *row = reinterpret_cast<SmT *>(fold_smem);
    const uint32_t NB = 1u << (2 * k);
    const uint32_t file = blockIdx.x;
    const uint32_t row0 = file_row[file], nrows = file_row[file + 1] - row0;
    const unsigned long long *g = g_fwd + (size_t)row0 * NB;
    const size_t orow = (size_t)file * (size_t)V;
    __shared__ unsigned long long red[FOLD_THREADS / 32];
    __shared__ unsigned long long s_total;
    if (NB >= 2) {
#pragma unroll 4
        for (uint32_t i = threadIdx.x; i < NB / 2; i += FOLD_THREADS) {
            ulonglong2 a = reinterpret_cast<const ulonglong2 *>(g)[i];
            for (uint32_t r = 1; r < nrows; r++) {
                const ulonglong2 v = reinterpret_cast<const ulonglong2 *>(g + (size_t)r * NB)[i];
                a.x += v.x;
                a.y += v.y;
            }
            row[2 * i] = (SmT)a.x;
            row[2 * i + 1] = (SmT)a.y;
        }
    } else {
        if (threadIdx.x == 0) { unsigned long long a = 0; for (uint32_t r = 0; r < nrows; r++) a += g[(size_t)r * NB]; row[0] = (SmT)a; }
    }
    __syncthreads();
    // rows written by the line kernel (k = 7, file_P != 0) hold the 7-mers with their digits reversed
    const bool rev = file_P != nullptr && file_P[file] != 0;
    unsigned long long c[FOLD_ITEMS];
    unsigned long long local = 0;
#pragma unroll
    for (int it = 0; it < FOLD_ITEMS; it++) {
        const long long i = (long long)threadIdx.x + (long long)it * FOLD_THREADS;
        c[it] = 0;
        if (i < V) {
            const uint32_t m = canon[i];
            const uint32_t r = revcomp_std(m, k);
            const uint32_t g0 = std_to_gray(m), g1 = std_to_gray(r);
            unsigned long long v = row[rev ? digit_rev7(g0) : g0];
            if (r != m) v += row[rev ? digit_rev7(g1) : g1];
            c[it] = v;
            if (counts) counts[orow + i] = v;
            local += v;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(FULL, local, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x < 32) {
        unsigned long long v = (threadIdx.x < FOLD_THREADS / 32) ? red[threadIdx.x] : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
        if (threadIdx.x == 0) { s_total = v; if (totals) totals[file] = v; }
    }
    __syncthreads();
    if (!freq && !feat) return;
    const bool pc = flags & 1u, raw = flags & 2u;
    const double denom = (double)s_total + (pc ? 0.5 * (double)V : 0.0);
#pragma unroll
    for (int it = 0; it < FOLD_ITEMS; it++) {
        const long long i = (long long)threadIdx.x + (long long)it * FOLD_THREADS;
        if (i < V) {
            double v = (double)c[it] + (pc ? 0.5 : 0.0);
            if (!raw) v = v / denom;   // IEEE fp64 division: correctly rounded, bit-exact with numpy
            if (freq) freq[orow + i] = v;
            if (feat) feat[orow + i] = (float)(v * 1e4);   // train_classifier_model.py:149,323
        }
    }
}

template <typename FwdT>
__global__ void __launch_bounds__(1024)
fold_normalize_kernel(const FwdT *__restrict__ g_fwd, const uint32_t *__restrict__ canon, int k, long long V,
                      uint32_t flags, uint32_t file_base, const uint32_t *__restrict__ file_P,
                      const uint32_t *__restrict__ file_row, unsigned long long *__restrict__ counts,
                      double *__restrict__ freq, float *__restrict__ feat,
                      unsigned long long *__restrict__ totals) {
    const size_t NB = (size_t)1 << (2 * k);
    const uint32_t file = blockIdx.x;
    // rows of this file: one, or (k = 7 line kernel) one per CTA that held a piece of it
    const uint32_t row0 = file_row ? file_row[file + file_base] : file;
    const uint32_t nrows = file_row ? file_row[file + file_base + 1] - row0 : 1u;
    const FwdT *g = g_fwd + (size_t)row0 * NB;
    const size_t orow = (size_t)(file + file_base) * (size_t)V;
    __shared__ unsigned long long red[32];
    __shared__ unsigned long long s_total;
    // rows written by the line kernel (k = 7, file_P != 0) hold the 7-mers with their digits reversed
    const bool rev = file_P != nullptr && file_P[file + file_base] != 0;
    auto row_index = [&](uint32_t m) -> uint32_t {
        const uint32_t gcode = std_to_gray(m);
        return rev ? digit_rev7(gcode) : gcode;
    };
    auto canon_count = [&](long long i) -> unsigned long long {
        const uint32_t m = canon[i];
        const uint32_t r = revcomp_std(m, k);
        const uint32_t i0 = row_index(m), i1 = row_index(r);
        unsigned long long c = 0;
        for (uint32_t rr = 0; rr < nrows; rr++) {
            c += (unsigned long long)g[(size_t)rr * NB + i0];
            if (r != m) c += (unsigned long long)g[(size_t)rr * NB + i1];
        }
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


// Large k, few files: the fold of one file cut into gridDim.y slices of the vocabulary (one CTA per file leaves most
// SMs idle when 4^k is large: k = 12 has 8.4 M canonical k-mers per file).  Two launches: canonical counts + per-file
// total (u64 atomics: exact), then the normalisation.  Rows are the k >= 8 layout: one u32 row per file.
__device__ __forceinline__ unsigned long long fold_canon_count(const uint32_t *__restrict__ g, uint32_t m, int k) {
    const uint32_t r = revcomp_std(m, k);
    unsigned long long c = g[std_to_gray(m)];
    if (r != m) c += g[std_to_gray(r)];
    return c;
}
__global__ void __launch_bounds__(1024)
fold_counts_sliced_kernel(const uint32_t *__restrict__ g_fwd, const uint32_t *__restrict__ canon, int k, long long V, uint32_t file_base,
                          unsigned long long *__restrict__ counts, unsigned long long *__restrict__ tot_ws) {
    const size_t NB = (size_t)1 << (2 * k);
    const uint32_t file = blockIdx.x;
    const uint32_t *g = g_fwd + (size_t)file * NB;
    const size_t orow = (size_t)(file + file_base) * (size_t)V;
    const long long per = (V + gridDim.y - 1) / gridDim.y;
    const long long i0 = (long long)blockIdx.y * per, i1 = i0 + per < V ? i0 + per : V;
    unsigned long long local = 0;
    for (long long i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
        const unsigned long long c = fold_canon_count(g, canon[i], k);
        if (counts) counts[orow + i] = c;
        local += c;
    }
    __shared__ unsigned long long red[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(FULL, local, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x < 32) {
        unsigned long long v = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
        if (threadIdx.x == 0 && v) atomicAdd(tot_ws + file, v);
    }
}
__global__ void __launch_bounds__(1024)
fold_norm_sliced_kernel(const uint32_t *__restrict__ g_fwd, const uint32_t *__restrict__ canon, int k, long long V, uint32_t flags,
                        uint32_t file_base, const unsigned long long *__restrict__ counts, double *__restrict__ freq,
                        float *__restrict__ feat, unsigned long long *__restrict__ totals, const unsigned long long *__restrict__ tot_ws) {
    const size_t NB = (size_t)1 << (2 * k);
    const uint32_t file = blockIdx.x;
    const uint32_t *g = g_fwd + (size_t)file * NB;
    const size_t orow = (size_t)(file + file_base) * (size_t)V;
    const unsigned long long total = tot_ws[file];
    if (totals && blockIdx.y == 0 && threadIdx.x == 0) totals[file + file_base] = total;
    if (!freq && !feat) return;
    const bool pc = flags & 1u, raw = flags & 2u;
    const double denom = (double)total + (pc ? 0.5 * (double)V : 0.0);
    const long long per = (V + gridDim.y - 1) / gridDim.y;
    const long long i0 = (long long)blockIdx.y * per, i1 = i0 + per < V ? i0 + per : V;
    for (long long i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
        const unsigned long long c = counts ? counts[orow + i] : fold_canon_count(g, canon[i], k);
        double v = (double)c + (pc ? 0.5 : 0.0);
        if (!raw) v = v / denom;   // IEEE fp64 division: correctly rounded, bit-exact with numpy
        if (freq) freq[orow + i] = v;
        if (feat) feat[orow + i] = (float)(v * 1e4);   // train_classifier_model.py:149,323
    }
}


// The same first launch with both gathers coalesced (k >= 8).  A k-mer in the sorted alphabet is cut into its first three
// bases a, the middle, and its last three bases b; a CTA takes one (file, middle) tile = 64 x 64 k-mers.  Forward bins of
// consecutive b lie in one 64-entry block of the row; the reverse complement of (a, mid, b) is (rc b, rc mid, rc a), so
// the reverse-complement bins of consecutive a lie in one 64-entry block as well: they are read with a fastest and
// turned through shared memory.  rank[m] = column of canonical m, 0xFFFFFFFF for the other strand (table per k).
constexpr int FOLD_T = 3;
__global__ void __launch_bounds__(1024)
fold_counts_tiled_kernel(const uint32_t *__restrict__ g_fwd, const uint32_t *__restrict__ rank, int k, long long V, uint32_t file_base,
                         unsigned long long *__restrict__ counts, unsigned long long *__restrict__ tot_ws) {
    constexpr int E = 1 << (2 * FOLD_T);   // 64
    __shared__ uint32_t R[E][E + 1];
    __shared__ unsigned long long red[32];
    const size_t NB = (size_t)1 << (2 * k);
    const uint32_t file = blockIdx.x;
    const uint32_t mid = blockIdx.y;                  // the k - 6 middle bases
    const uint32_t *g = g_fwd + (size_t)file * NB;
    const size_t orow = (size_t)(file + file_base) * (size_t)V;
    const int midbits = 2 * (k - 2 * FOLD_T);
    const uint32_t lo = threadIdx.x & (E - 1), hi0 = threadIdx.x >> (2 * FOLD_T);   // 1024 threads: 16 rows of 64 per step
    // reverse-complement bins, a fastest: thread (b = hi, a = lo) reads the bin of rc(a, mid, b) and stores it at R[a][b]
#pragma unroll
    for (int step = 0; step < E / 16; step++) {
        const uint32_t b = hi0 + 16u * step, a = lo;
        const uint32_t m = (a << (midbits + 2 * FOLD_T)) | (mid << (2 * FOLD_T)) | b;
        R[a][b] = g[std_to_gray(revcomp_std(m, k))];
    }
    __syncthreads();
    unsigned long long local = 0;
#pragma unroll
    for (int step = 0; step < E / 16; step++) {
        const uint32_t a = hi0 + 16u * step, b = lo;
        const uint32_t m = (a << (midbits + 2 * FOLD_T)) | (mid << (2 * FOLD_T)) | b;
        const uint32_t col = rank[m];
        if (col != 0xFFFFFFFFu) {
            unsigned long long c = g[std_to_gray(m)];
            if (revcomp_std(m, k) != m) c += R[a][b];
            if (counts) counts[orow + col] = c;
            local += c;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(FULL, local, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x < 32) {
        unsigned long long v = red[threadIdx.x];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
        if (threadIdx.x == 0 && v) atomicAdd(tot_ws + file, v);
    }
}

}  // namespace kf
