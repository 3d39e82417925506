// kf_sparse.cuh -- sparse (sort-and-run-length) k-mer counting for large k (sm_100a).
//
// Replaces `jellyfish count -m k -C` + `jellyfish dump -c` (kf2vec/main.py:135-145, get_kmers; main.py:308-319) where a
// dense row of 4^k bins is no longer sensible (k = 12: 64 MB per file; k up to 31 is what the reference's -k accepts,
// main.py:81-82): the result is the list of OBSERVED canonical k-mers of every file with their counts, ascending by code
// (A0 C1 G2 T3, first base most significant -- the vocabulary order of the dense path), which is what Jellyfish dumps
// (in hash order).
//
// k = 9 .. 12 take the 16-bit-bucket pipeline further down (sparse_wc_scatter_kernel, sparse16_*): buckets of 65,536 codes,
// 16-bit keys, write-combined partition -- 3 to 5 times the rate of what follows.  Everything else (k = 6 .. 8, k >= 13):
// pipeline over one batch of files (MSD radix partition by the code's leading 12 bits, then per bucket):
//   1. sparse_tile_kernel<0>      text -> canonical codes -> per-(tile, bucket) counts (shared-memory histogram per tile)
//   2. sparse_tile_scan_kernel    per file: where every (tile, bucket) run goes inside the file's key region
//   3. sparse_tile_kernel<1>      the same parse again (the tile comes from L2); every code is stored at its place
//                                 (cursors in shared memory: no global atomics)
//   k <= 12 (a bucket's codes differ in <= 12 bits): histogram instead of sort
//   4. sparse_bucket_distinct_kernel   one warp per bucket: bit map -> number of distinct codes
//   5. sparse_scan_distinct_kernel     per file: where every bucket's entries go
//   6. sparse_bucket_emit_kernel       one warp per bucket: 2^R bins in shared memory -> (code, count) ascending
//   k >= 13: sort
//   4. sparse_sort_kernel         item = (file, run of buckets): each bucket sorted in shared memory (bitonic; in global
//                                 memory in place when it does not fit), distinct codes counted
//   5. sparse_scan_items_kernel   exclusive scan of the distinct counts: where every item's output goes
//   6. sparse_emit_kernel         run-length encode the item's (now globally sorted) key range into (code, count)
// The parser is the dense path's (fasta_process_range): a k-mer is owned by the 16-byte lane that holds its first base,
// so both passes enumerate exactly the same k-mers.  k <= 16: the lane's 32-base window yields forward and
// reverse-complement codes with two shifts each; k >= 17: every lane runs the canonical byte walker.
#pragma once
#include "kf_kernels.cuh"

namespace kf {

constexpr int SP_BUCKET_BITS = 12;                  // leading bits of the 2k-bit canonical code that number the bucket
constexpr uint32_t SP_BUCKETS = 1u << SP_BUCKET_BITS;
constexpr int SP_MIN_K = 6;                         // (the code must have at least SP_BUCKET_BITS bits)
constexpr int SP_MAX_K = 31;

__device__ __forceinline__ uint64_t digit_reverse64(uint64_t x) {   // reverse the 32 two-bit digits
    x = __brevll(x);
    return ((x & 0xAAAAAAAAAAAAAAAAull) >> 1) | ((x & 0x5555555555555555ull) << 1);
}

// What a counted canonical code does: MODE 0 bumps the bucket histogram in shared memory, MODE 1 stores the code in its
// bucket of the file's key region.
template <int MODE, typename KT>
struct SparseEmit {
    uint32_t *hist;       // MODE 0: shared, SP_BUCKETS bins
    uint32_t *cursor;     // MODE 1: global, SP_BUCKETS cursors of the current file (offsets inside its key region)
    KT *keys;             // MODE 1: the current file's key region
    uint32_t shift;       // 2k - SP_BUCKET_BITS
    __device__ __forceinline__ void operator()(uint64_t canon) const {
        const uint32_t b = (uint32_t)(canon >> shift);
        if (MODE == 0) atomicAdd(hist + b, 1u);
        else keys[atomicAdd(cursor + b, 1u)] = (KT)canon;
    }
};

// Canonical byte walker, run-time k (1..31): counts every k-mer whose first base lies in [p0, p1); (in_hdr, at_ls) is the
// line state at p0.  Same contract as fasta_walk_lane.  Codes in the sorted alphabet A0 C1 G2 T3.
template <class Src, class Emit>
__device__ KF_NOINLINE void fasta_walk_lane_canon(const Src src, uint64_t p0, uint64_t p1, bool in_hdr, bool at_ls, int k, Emit emit) {
    const uint64_t mask = (1ull << (2 * k)) - 1ull;
    const int rsh = 2 * (k - 1);
    uint64_t F = 0, R = 0;
    int run = 0, owned = 0;
    uint64_t p = p0;
    for (;;) {
        const bool own = p < p1;
        if (!own && (owned == 0 || run - k + 1 >= owned)) break;
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
        const uint32_t gcode = (c >> 1) & 3u;
        const uint64_t sc = (uint64_t)(gcode ^ (gcode >> 1));   // A0 C1 T2 G3 -> A0 C1 G2 T3
        F = ((F << 2) | sc) & mask;
        R = (R >> 2) | ((3ull - sc) << rsh);
        run++;
        if (own) owned++;
        if (run >= k && run - k < owned) emit(F < R ? F : R);
    }
}

template <int MODE, typename KT>
struct SparseSink {
    SparseEmit<MODE, KT> e;
    int k;
    // fast lanes (k <= 16): the k-mers that start at bases 0..n-1 of the 32-base window hi:lo (gray codes, first base in
    // hi bits 31:30)
    __device__ __forceinline__ void window(uint32_t hi, uint32_t lo, uint32_t n) const {
        const uint32_t hs = hi ^ ((hi >> 1) & 0x55555555u), ls = lo ^ ((lo >> 1) & 0x55555555u);   // sorted alphabet
        const uint64_t w = ((uint64_t)hs << 32) | ls;
        const uint64_t rcw = digit_reverse64(~w);       // digit d (from the top) = complement of base 31 - d
        const uint64_t mask = (1ull << (2 * k)) - 1ull;
        const int s0 = 64 - 2 * k;
#pragma unroll
        for (int j = 0; j < 16; j++) {
            if (j < 15 || n == 16) {
                const uint64_t F = (w >> (s0 - 2 * j)) & mask;
                const uint64_t R = (rcw >> (2 * j)) & mask;
                e(F < R ? F : R);
            }
        }
    }
    template <class Src>
    __device__ __forceinline__ void walk(const Src src, uint64_t p0, uint64_t p1, bool in_hdr, bool at_ls) const {
        fasta_walk_lane_canon(src, p0, p1, in_hdr, at_ls, k, e);
    }
    __device__ __forceinline__ void operator()(uint32_t) const {}   // (never reached: walk() takes the rare paths)
};
template <int MODE, typename KT> struct sink_takes_window<SparseSink<MODE, KT>> { static constexpr bool value = true; };
template <int MODE, typename KT> struct sink_walks_itself<SparseSink<MODE, KT>> { static constexpr bool value = true; };

// ------------------------------------------------------------------------------------------------
// Partition without global atomics (passes 1-3 of the pipeline in kf_sparse_host.inc)
// ------------------------------------------------------------------------------------------------
// A tile (<= 512 KiB of one file, the dense path's plan) is one CTA's unit: MODE 0 counts the tile's canonical codes per
// bucket in shared memory and writes the 4,096 counts to tile_hist[tile]; the scan below turns them into the place of
// every (tile, bucket) run inside the file's key region; MODE 1 loads those places as shared-memory cursors, parses the
// tile again (it comes from L2) and stores every code at atomicAdd(cursor[bucket], 1) -- a shared-memory atomic.  The first
// version took the cursor from global memory: one returning global atomic per k-mer, 7 Gbases/s.
template <int MODE, typename KT, bool WALK_ALL, int THREADS>
__global__ void __launch_bounds__(THREADS, 2)
sparse_tile_kernel(const uint8_t *__restrict__ arena, const Tile *__restrict__ tiles, const int *__restrict__ cta_begin, int k,
                   uint32_t file_base, uint32_t *__restrict__ tile_hist, KT *__restrict__ keys, const uint64_t *__restrict__ kbase,
                   int bucket_shift /* bucket = code >> bucket_shift; < 0: the leading SP_BUCKET_BITS bits */) {
    __shared__ uint32_t hist[SP_BUCKETS];
    constexpr int NWARPS = THREADS / 32;
    const int warp = threadIdx.x >> 5;
    SparseSink<MODE, KT> sink;
    sink.k = k;
    sink.e.hist = hist;
    sink.e.cursor = hist;
    sink.e.shift = bucket_shift >= 0 ? (uint32_t)bucket_shift : (uint32_t)(2 * k - SP_BUCKET_BITS);
    sink.e.keys = nullptr;
    const int t1 = cta_begin[blockIdx.x + 1];
    for (int t = cta_begin[blockIdx.x]; t < t1; ++t) {
        const Tile T = tiles[t];
        uint4 *th4 = reinterpret_cast<uint4 *>(tile_hist + (size_t)t * SP_BUCKETS);
        uint4 *h4 = reinterpret_cast<uint4 *>(hist);
        for (uint32_t i = threadIdx.x; i < SP_BUCKETS / 4; i += THREADS) h4[i] = MODE == 0 ? make_uint4(0u, 0u, 0u, 0u) : th4[i];
        __syncthreads();
        if (MODE == 1) sink.e.keys = keys + kbase[T.file - file_base];
        const uint32_t cpw = (T.n_chunks + NWARPS - 1) / NWARPS;
        const uint32_t c0 = T.first_chunk + (uint32_t)warp * cpw;
        const uint32_t cend = T.first_chunk + T.n_chunks;
        const uint32_t c1 = (c0 + cpw < cend) ? c0 + cpw : cend;
        if (c0 < c1) fasta_process_range<12, WALK_ALL, 3>(GlobalSrc{arena}, c0, c1, T.file_chunk0, sink);
        __syncthreads();
        if (MODE == 0)
            for (uint32_t i = threadIdx.x; i < SP_BUCKETS / 4; i += THREADS) th4[i] = h4[i];
        __syncthreads();
    }
}

// Pass 2: one CTA per file.  tile_hist [tiles][SP_BUCKETS]: counts in, places out (offset of the (tile, bucket) run inside
// the file's key region); boff [nf][SP_BUCKETS + 1] = exclusive scan of the file's bucket totals (last = all keys).
__global__ void __launch_bounds__(1024)
sparse_tile_scan_kernel(uint32_t *__restrict__ tile_hist, const int *__restrict__ file_t0, uint32_t *__restrict__ boff,
                        unsigned long long *__restrict__ totals, uint32_t file_base,
                        uint32_t pad /* 0, or 7: every run starts at a multiple of 8 keys */, uint32_t *__restrict__ tile_place /* null: in place */) {
    static_assert(SP_BUCKETS == 4096, "four buckets per thread");
    __shared__ uint32_t wsum[32];
    __shared__ unsigned long long s_true;
    if (threadIdx.x == 0) s_true = 0ull;
    if (tile_place == nullptr) tile_place = tile_hist;
    const uint32_t f = blockIdx.x;
    const int ta = file_t0[f], tb = file_t0[f + 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint4 run = make_uint4(0u, 0u, 0u, 0u);
    unsigned long long tsum = 0;   // the true number of keys (pad or not)
    for (int t = ta; t < tb; t++) {   // per bucket: counts -> running offsets over the file's tiles
        const uint4 c = reinterpret_cast<const uint4 *>(tile_hist + (size_t)t * SP_BUCKETS)[threadIdx.x];
        reinterpret_cast<uint4 *>(tile_place + (size_t)t * SP_BUCKETS)[threadIdx.x] = run;
        run.x += (c.x + pad) & ~pad; run.y += (c.y + pad) & ~pad; run.z += (c.z + pad) & ~pad; run.w += (c.w + pad) & ~pad;
        tsum += (unsigned long long)c.x + c.y + c.z + c.w;
    }
    if (pad) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tsum += __shfl_xor_sync(FULL, tsum, o);
        __syncthreads();
        if ((threadIdx.x & 31) == 0 && tsum) atomicAdd(&s_true, tsum);
    }
    const uint4 v = run;   // the file's bucket totals
    const uint32_t mine = v.x + v.y + v.z + v.w;
    uint32_t inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = wsum[lane], wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(FULL, wi, o); if (lane >= o) wi += t; }
        wsum[lane] = wi - w;
    }
    __syncthreads();
    const uint32_t ex = wsum[warp] + inc - mine;
    const uint4 o4 = make_uint4(ex, ex + v.x, ex + v.x + v.y, ex + v.x + v.y + v.z);
    uint32_t *bo = boff + (size_t)f * (SP_BUCKETS + 1);
    bo[4 * threadIdx.x] = o4.x; bo[4 * threadIdx.x + 1] = o4.y; bo[4 * threadIdx.x + 2] = o4.z; bo[4 * threadIdx.x + 3] = o4.w;
    if (threadIdx.x == 1023) {
        bo[SP_BUCKETS] = ex + mine;
        if (totals) totals[file_base + f] = pad ? s_true : (unsigned long long)(ex + mine);   // (s_true: complete since the barriers above)
    }
    for (int t = ta; t < tb; t++) {   // + the bucket's start
        uint4 *p = reinterpret_cast<uint4 *>(tile_place + (size_t)t * SP_BUCKETS) + threadIdx.x;
        uint4 c = *p;
        c.x += o4.x; c.y += o4.y; c.z += o4.z; c.w += o4.w;
        *p = c;
    }
}

// ------------------------------------------------------------------------------------------------
// Buckets of few remaining bits (2k - 12 <= 12, i.e. k <= 12): histogram instead of sort
// ------------------------------------------------------------------------------------------------
// The codes of one bucket differ in their low R = 2k - 12 bits only: at most 4,096 values.  One warp per bucket:
//   sparse_bucket_distinct_kernel  a 2^R-bit map in shared memory (atomicOr) -> number of distinct codes of the bucket
//   sparse_scan_distinct_kernel    per file: exclusive scan of the 4,096 numbers -> where the bucket's entries go
//   sparse_bucket_emit_kernel      2^R u32 bins in shared memory per warp (atomicAdd), then the bins are walked 32 at a time
//                                  (ballot + popc give the place of every non-zero bin: ascending codes, coalesced stores)
constexpr int SP_HIST_MAX_R = 12;
constexpr int SP_HIST_WARPS = 8;      // warps (= buckets in flight) per CTA: 8 x 16 KB of bins

__global__ void __launch_bounds__(32 * SP_HIST_WARPS)
sparse_bucket_distinct_kernel(const uint32_t *__restrict__ keys, const uint64_t *__restrict__ kbase, const uint32_t *__restrict__ boff, int R,
                              uint32_t *__restrict__ nd /* [nf][SP_BUCKETS] */) {
    __shared__ uint32_t bitmap[SP_HIST_WARPS][(1 << SP_HIST_MAX_R) / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t item = blockIdx.x * SP_HIST_WARPS + warp;      // (file, bucket)
    const uint32_t f = item >> SP_BUCKET_BITS, b = item & (SP_BUCKETS - 1);
    const uint32_t nwords = (1u << R) / 32u ? (1u << R) / 32u : 1u;
    uint32_t *bm = bitmap[warp];
    for (uint32_t i = lane; i < nwords; i += 32) bm[i] = 0;
    KF_SYNCWARP();
    const uint32_t *bo = boff + (size_t)f * (SP_BUCKETS + 1);
    const uint32_t s = bo[b], n = bo[b + 1] - s;
    const uint32_t *gk = keys + kbase[f] + s;
    const uint32_t lowmask = (1u << R) - 1u;
    for (uint32_t i = lane; i < n; i += 32) {
        const uint32_t low = gk[i] & lowmask;
        atomicOr(bm + (low >> 5), 1u << (low & 31u));
    }
    KF_SYNCWARP();
    uint32_t c = 0;
    for (uint32_t i = lane; i < nwords; i += 32) c += (uint32_t)__popc(bm[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(FULL, c, o);
    if (lane == 0) nd[item] = c;
}

// per file: ooff [nf][SP_BUCKETS] = exclusive scan of nd over the file's buckets, nd_file [nf] = the file's distinct codes
__global__ void __launch_bounds__(1024)
sparse_scan_distinct_kernel(const uint32_t *__restrict__ nd, uint32_t *__restrict__ ooff, unsigned long long *__restrict__ nd_file) {
    __shared__ uint32_t wsum[32];
    const uint32_t f = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint4 v = reinterpret_cast<const uint4 *>(nd + (size_t)f * SP_BUCKETS)[threadIdx.x];
    const uint32_t mine = v.x + v.y + v.z + v.w;
    uint32_t inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = wsum[lane], wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(FULL, wi, o); if (lane >= o) wi += t; }
        wsum[lane] = wi - w;
    }
    __syncthreads();
    const uint32_t ex = wsum[warp] + inc - mine;
    reinterpret_cast<uint4 *>(ooff + (size_t)f * SP_BUCKETS)[threadIdx.x] = make_uint4(ex, ex + v.x, ex + v.x + v.y, ex + v.x + v.y + v.z);
    if (threadIdx.x == 1023) nd_file[f] = (unsigned long long)(ex + mine);
}

__global__ void __launch_bounds__(32 * SP_HIST_WARPS)
sparse_bucket_emit_kernel(const uint32_t *__restrict__ keys, const uint64_t *__restrict__ kbase, const uint32_t *__restrict__ boff, int R,
                          const uint32_t *__restrict__ ooff, const unsigned long long *__restrict__ out_base /* [nf]: first entry of the file */,
                          unsigned long long *__restrict__ codes_out, uint32_t *__restrict__ counts_out) {
    KF_DYN_SMEM(uint32_t, sp_bins);                               // SP_HIST_WARPS x 2^R u32
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t item = blockIdx.x * SP_HIST_WARPS + warp;      // (file, bucket)
    const uint32_t f = item >> SP_BUCKET_BITS, b = item & (SP_BUCKETS - 1);
    const uint32_t nbins = 1u << R;
    uint32_t *bins = sp_bins + (size_t)warp * nbins;
    const uint32_t *bo = boff + (size_t)f * (SP_BUCKETS + 1);
    const uint32_t s = bo[b], n = bo[b + 1] - s;
    if (n == 0) return;                                           // (the whole warp)
    for (uint32_t i = lane; i < nbins; i += 32) bins[i] = 0;
    KF_SYNCWARP();
    const uint32_t *gk = keys + kbase[f] + s;
    const uint32_t lowmask = nbins - 1u;
    for (uint32_t i = lane; i < n; i += 32) atomicAdd(bins + (gk[i] & lowmask), 1u);
    KF_SYNCWARP();
    unsigned long long pos = out_base[f] + (unsigned long long)ooff[item];
    const unsigned long long hi = (unsigned long long)b << R;
    for (uint32_t i0 = 0; i0 < nbins; i0 += 32) {
        const uint32_t i = i0 + (uint32_t)lane;
        const uint32_t c = i < nbins ? bins[i] : 0u;
        const unsigned m = __ballot_sync(FULL, c != 0u);
        if (c) {
            const unsigned long long o = pos + (unsigned long long)__popc(m & ((1u << lane) - 1u));
            codes_out[o] = hi | (unsigned long long)i;
            counts_out[o] = c;
        }
        pos += (unsigned long long)__popc(m);
    }
}

// ------------------------------------------------------------------------------------------------
// 16 < 2k <= 24 (k = 9 .. 12): buckets of 65,536 codes, write-combined partition, histogram per (file, bucket)
// ------------------------------------------------------------------------------------------------
// What bounded the 4,096-bucket partition above was not its atomics but its stores: a warp's 32 codes go to 32 different
// runs, i.e. 32 write requests of 4 bytes, and an SM hands the L2 about one request every four clocks whatever its size
// (profiles/r02_v2_ncu_sparse.txt: 61 G stores/s, issue slots 6 % busy, DRAM writes 3.5 x the keys).  Here a code's leading
// 2k - 16 bits number its bucket (<= 256 buckets) and only its low 16 bits are kept:
//   sparse_wc_scatter_kernel   every WARP owns 16 two-byte slots per bucket in shared memory; a code goes to its bucket's
//                              slots (shared-memory atomic), full groups of 8 leave as ONE 16-byte store -- 8 codes per
//                              write request, and half the bytes.  A (tile, bucket) run starts at a multiple of 8 keys
//                              (the scan pads); what is left in the slots at the tile's end, and whatever arrives at a full
//                              bucket, is stored code by code from the run's END downwards: blocks and single codes meet
//                              without a hole, the padding lies behind the run's count.
//   sparse16_distinct_kernel   one CTA per (file, bucket): 65,536-bit map -> number of distinct codes
//   sparse16_emit_kernel       one CTA per (file, bucket): two halves of 32,768 u32 bins (128 KB), each warp walks its
//                              2,048 bins in order (count, scan over the warps, write): ascending (code, count) entries
constexpr int SP16_SLOTS = 16;
constexpr int SP16_THREADS_SCATTER = 256;   // 8 warps x 256 buckets x 16 slots x 2 bytes = 64 KB (+ fills and cursors): two CTAs per SM
constexpr int SP16_THREADS_EMIT = 512;
constexpr uint32_t SP16_MAX_BUCKETS = 256;
constexpr size_t sp16_scatter_smem() {
    return (size_t)(SP16_THREADS_SCATTER / 32) * SP16_MAX_BUCKETS * (SP16_SLOTS * sizeof(uint16_t) + sizeof(uint32_t)) + 2 * SP16_MAX_BUCKETS * sizeof(uint32_t);
}

struct SparseEmitWC {
    uint16_t *wbuf;     // this warp's [buckets][SP16_SLOTS]
    uint32_t *fill;     // this warp's [buckets]: codes put into the slots since the bucket was last drained (may exceed the slots)
    uint32_t *back;     // CTA: [buckets] end of the run's unwritten part (single codes are stored below it)
    uint16_t *keys;     // the current file's key region
    __device__ __forceinline__ void operator()(uint64_t canon) const {
        const uint32_t b = (uint32_t)(canon >> 16);
        const uint32_t slot = atomicAdd(fill + b, 1u);
        if (slot < (uint32_t)SP16_SLOTS) wbuf[b * SP16_SLOTS + slot] = (uint16_t)canon;
        else keys[atomicSub(back + b, 1u) - 1u] = (uint16_t)canon;
    }
};
struct SparseWcSink {
    SparseEmitWC e;
    uint32_t *front;    // CTA: [buckets] where the run's next block of 8 goes
    uint32_t nbuckets;
    int k;
    __device__ __forceinline__ void window(uint32_t hi, uint32_t lo, uint32_t n) const {   // (SparseSink::window)
        const uint32_t hs = hi ^ ((hi >> 1) & 0x55555555u), ls = lo ^ ((lo >> 1) & 0x55555555u);
        const uint64_t w = ((uint64_t)hs << 32) | ls;
        const uint64_t rcw = digit_reverse64(~w);
        const uint64_t mask = (1ull << (2 * k)) - 1ull;
        const int s0 = 64 - 2 * k;
#pragma unroll
        for (int j = 0; j < 16; j++) {
            if (j < 15 || n == 16) {
                const uint64_t F = (w >> (s0 - 2 * j)) & mask;
                const uint64_t R = (rcw >> (2 * j)) & mask;
                e(F < R ? F : R);
            }
        }
    }
    template <class Src>
    __device__ __forceinline__ void walk(const Src src, uint64_t p0, uint64_t p1, bool in_hdr, bool at_ls) const {
        fasta_walk_lane_canon(src, p0, p1, in_hdr, at_ls, k, e);
    }
    // All lanes, converged (fasta_process_range calls record() for every lane of every chunk before the chunk's k-mers are
    // counted; the kernel calls it once more after the range): buckets that hold 8 codes or more give their first 8 (or 16)
    // away as 16-byte stores, the rest moves to the front of the slots.  Lane l looks after the buckets l, l + 32, ...
    __device__ __forceinline__ void drain(bool all) const {
        KF_SYNCWARP();
        const int lane = threadIdx.x & 31;
        for (uint32_t b = (uint32_t)lane; b < nbuckets; b += 32) {
            uint32_t n = e.fill[b];
            if (n > (uint32_t)SP16_SLOTS) n = SP16_SLOTS;
            if (n < 8u && !(all && n > 0u)) continue;
            uint4 *slots = reinterpret_cast<uint4 *>(e.wbuf + b * SP16_SLOTS);
            uint32_t done = 0;
            while (n - done >= 8u) {
                const uint32_t at = atomicAdd(front + b, 8u);
                *reinterpret_cast<uint4 *>(e.keys + at) = slots[done >> 3];
                done += 8u;
            }
            uint32_t rest = n - done;   // < 8
            if (all) {
                for (uint32_t i = 0; i < rest; i++) e.keys[atomicSub(e.back + b, 1u) - 1u] = e.wbuf[b * SP16_SLOTS + done + i];
                rest = 0;
            } else if (done == 8u && rest) {
                slots[0] = slots[1];
            }
            e.fill[b] = rest;
        }
        KF_SYNCWARP();
    }
    __device__ __forceinline__ void record(size_t, bool, uint32_t, uint32_t, uint32_t) const { drain(false); }
    __device__ __forceinline__ void operator()(uint32_t) const {}
};
template <> struct sink_takes_window<SparseWcSink> { static constexpr bool value = true; };
template <> struct sink_walks_itself<SparseWcSink> { static constexpr bool value = true; };
template <> struct sink_records_lanes<SparseWcSink> { static constexpr bool value = true; };

// tile_cnt / tile_place [tiles][SP_BUCKETS]: a (tile, bucket) run's number of codes and its place in the file's key region
// (multiple of 8); keys16: the batch's key workspace, file f at kbase[f] (multiple of 8).
template <int THREADS>
__global__ void __launch_bounds__(THREADS, 2)
sparse_wc_scatter_kernel(const uint8_t *__restrict__ arena, const Tile *__restrict__ tiles, const int *__restrict__ cta_begin, int k,
                         uint32_t file_base, const uint32_t *__restrict__ tile_cnt, const uint32_t *__restrict__ tile_place,
                         uint16_t *__restrict__ keys16, const uint64_t *__restrict__ kbase) {
    constexpr int NWARPS = THREADS / 32;
    KF_DYN_SMEM(uint32_t, wc_smem);
    const uint32_t NBK = 1u << (2 * k - 16);
    uint32_t *front = wc_smem, *back = front + SP16_MAX_BUCKETS;
    uint32_t *fill_all = back + SP16_MAX_BUCKETS;
    uint16_t *wbuf_all = reinterpret_cast<uint16_t *>(fill_all + NWARPS * SP16_MAX_BUCKETS);
    const int warp = threadIdx.x >> 5;
    SparseWcSink sink;
    sink.k = k;
    sink.nbuckets = NBK;
    sink.front = front;
    sink.e.back = back;
    sink.e.fill = fill_all + warp * SP16_MAX_BUCKETS;
    sink.e.wbuf = wbuf_all + (size_t)warp * SP16_MAX_BUCKETS * SP16_SLOTS;
    sink.e.keys = nullptr;
    for (uint32_t i = threadIdx.x; i < NWARPS * SP16_MAX_BUCKETS; i += THREADS) fill_all[i] = 0;
    const int t1 = cta_begin[blockIdx.x + 1];
    for (int t = cta_begin[blockIdx.x]; t < t1; ++t) {
        const Tile T = tiles[t];
        for (uint32_t b = threadIdx.x; b < NBK; b += THREADS) {
            const uint32_t pl = tile_place[(size_t)t * SP_BUCKETS + b];
            front[b] = pl;
            back[b] = pl + tile_cnt[(size_t)t * SP_BUCKETS + b];
        }
        __syncthreads();
        sink.e.keys = keys16 + kbase[T.file - file_base];
        const uint32_t cpw = (T.n_chunks + NWARPS - 1) / NWARPS;
        const uint32_t c0 = T.first_chunk + (uint32_t)warp * cpw;
        const uint32_t cend = T.first_chunk + T.n_chunks;
        const uint32_t c1 = (c0 + cpw < cend) ? c0 + cpw : cend;
        if (c0 < c1) fasta_process_range<12, false, 3>(GlobalSrc{arena}, c0, c1, T.file_chunk0, sink);
        sink.drain(true);
        __syncthreads();
    }
}

// keys of (file f, bucket b): the runs (tile, b) of the file's tiles
__global__ void __launch_bounds__(256)
sparse16_distinct_kernel(const uint16_t *__restrict__ keys16, const uint64_t *__restrict__ kbase, const int *__restrict__ file_t0,
                         const uint32_t *__restrict__ tile_cnt, const uint32_t *__restrict__ tile_place, uint32_t nbuckets,
                         uint32_t *__restrict__ nd /* [nf][SP_BUCKETS], zeroed */) {
    __shared__ uint32_t bitmap[2048];
    __shared__ uint32_t s_n[2];   // distinct codes below / from 32,768 (the emit kernel takes the two halves in separate CTAs)
    const uint32_t f = blockIdx.x / nbuckets, b = blockIdx.x - f * nbuckets;
    for (uint32_t i = threadIdx.x; i < 2048; i += blockDim.x) bitmap[i] = 0;
    if (threadIdx.x < 2) s_n[threadIdx.x] = 0;
    __syncthreads();
    const uint16_t *fk = keys16 + kbase[f];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (int t = file_t0[f] + warp; t < file_t0[f + 1]; t += nwarps) {   // a warp per run, eight keys per 16-byte load
        const uint32_t n = tile_cnt[(size_t)t * SP_BUCKETS + b];
        const uint16_t *gk = fk + tile_place[(size_t)t * SP_BUCKETS + b];
        const uint32_t n8 = n & ~7u;
        for (uint32_t i = 8u * (uint32_t)lane; i < n8; i += 256u) {
            const uint4 q = *reinterpret_cast<const uint4 *>(gk + i);
            const uint32_t w4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint32_t v0 = w4[j] & 0xFFFFu, v1 = w4[j] >> 16;
                atomicOr(bitmap + (v0 >> 5), 1u << (v0 & 31u));
                atomicOr(bitmap + (v1 >> 5), 1u << (v1 & 31u));
            }
        }
        if (n8 + (uint32_t)lane < n) {
            const uint32_t v = gk[n8 + lane];
            atomicOr(bitmap + (v >> 5), 1u << (v & 31u));
        }
    }
    __syncthreads();
    uint32_t c0 = 0, c1 = 0;
    for (uint32_t i = threadIdx.x; i < 1024; i += blockDim.x) { c0 += (uint32_t)__popc(bitmap[i]); c1 += (uint32_t)__popc(bitmap[1024 + i]); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { c0 += __shfl_xor_sync(FULL, c0, o); c1 += __shfl_xor_sync(FULL, c1, o); }
    if ((threadIdx.x & 31) == 0) { if (c0) atomicAdd(&s_n[0], c0); if (c1) atomicAdd(&s_n[1], c1); }
    __syncthreads();
    if (threadIdx.x < 2) nd[(size_t)f * SP_BUCKETS + 2 * b + threadIdx.x] = s_n[threadIdx.x];
}

// One CTA per (file, bucket, half of the bucket's code range): 32,768 codes in 16,384 words of two u16 counters (code c of
// the half: word c >> 1, half word c & 1) = 64 KB, three CTAs per SM.  One pass over the bucket's keys (a warp per
// (tile, bucket) run, eight keys per 16-byte load; the other half's keys are skipped).  A counter that reaches 65,536
// carries or wraps: the sum of all counters then differs from the number of keys counted and the half is done again
// exactly, with u32 bins in two quarters (poly-A stretches; tests force it).  Then every warp walks its words in order,
// 128 per round (four words = eight codes per lane): count the non-zero counters, scan over the warps, write ascending
// (code, count) entries.  ooff / nd are indexed by 2 * bucket + half.
__device__ __forceinline__ uint32_t sp16_nonzero_halves(uint32_t w) {   // how many of the two u16 halves of w are non-zero
    return (uint32_t)__popc((((w & 0x7FFF7FFFu) + 0x7FFF7FFFu) | w) & 0x80008000u);
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS, 3)
sparse16_emit_kernel(const uint16_t *__restrict__ keys16, const uint64_t *__restrict__ kbase, const int *__restrict__ file_t0,
                     const uint32_t *__restrict__ tile_cnt, const uint32_t *__restrict__ tile_place, uint32_t nbuckets,
                     const uint32_t *__restrict__ ooff /* [nf][SP_BUCKETS] */, const unsigned long long *__restrict__ out_base /* [nf] */,
                     unsigned long long *__restrict__ codes_out, uint32_t *__restrict__ counts_out) {
    constexpr int NWARPS = THREADS / 32;
    constexpr uint32_t NWORDS = 16384;
    constexpr uint32_t PER_WARP = NWORDS / NWARPS;   // words a warp walks
    static_assert(PER_WARP % 128 == 0, "whole rounds of 128 words");
    KF_DYN_SMEM(uint32_t, bins);                                  // NWORDS u32
    __shared__ uint32_t s_w[3 * NWARPS];                          // per warp: non-zero counters | sum of all counters | keys counted
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t item = blockIdx.x >> 1, half = blockIdx.x & 1u;
    const uint32_t f = item / nbuckets, b = item - f * nbuckets;
    const uint16_t *fk = keys16 + kbase[f];
    const int ta = file_t0[f], tb = file_t0[f + 1];
    const unsigned long long pos = out_base[f] + (unsigned long long)ooff[(size_t)f * SP_BUCKETS + 2 * b + half];
    uint32_t total = 0;
    for (int t = ta; t < tb; t++) total += tile_cnt[(size_t)t * SP_BUCKETS + b];
    if (total == 0) return;   // (the whole CTA)
    uint4 *b4 = reinterpret_cast<uint4 *>(bins);
    for (uint32_t i = threadIdx.x; i < NWORDS / 4; i += THREADS) b4[i] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
    uint32_t nkeys = 0;   // keys of my half this thread has counted
    auto count = [&](uint32_t v) {
        if ((v >> 15) == half) { atomicAdd(bins + ((v & 0x7FFFu) >> 1), 1u << ((v & 1u) << 4)); nkeys++; }
    };
    for (int t = ta + warp; t < tb; t += NWARPS) {   // a warp per run; the run starts at a multiple of 8 keys
        const uint32_t n = tile_cnt[(size_t)t * SP_BUCKETS + b];
        const uint16_t *gk = fk + tile_place[(size_t)t * SP_BUCKETS + b];
        const uint32_t n8 = n & ~7u;
        for (uint32_t i = 8u * (uint32_t)lane; i < n8; i += 256u) {
            const uint4 q = *reinterpret_cast<const uint4 *>(gk + i);
            count(q.x & 0xFFFFu); count(q.x >> 16); count(q.y & 0xFFFFu); count(q.y >> 16);
            count(q.z & 0xFFFFu); count(q.z >> 16); count(q.w & 0xFFFFu); count(q.w >> 16);
        }
        if (n8 + (uint32_t)lane < n) count((uint32_t)gk[n8 + lane]);
    }
    __syncthreads();
    const uint4 *wb = reinterpret_cast<const uint4 *>(bins + (size_t)warp * PER_WARP);
    uint32_t nz = 0, sum = 0;
    for (uint32_t r = 0; r < PER_WARP / 128; r++) {
        const uint4 v = wb[r * 32 + lane];
        const uint32_t c4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; j++) {
            nz += sp16_nonzero_halves(c4[j]);
            sum += (c4[j] & 0xFFFFu) + (c4[j] >> 16);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { nz += __shfl_xor_sync(FULL, nz, o); sum += __shfl_xor_sync(FULL, sum, o); nkeys += __shfl_xor_sync(FULL, nkeys, o); }
    if (lane == 0) { s_w[warp] = nz; s_w[NWARPS + warp] = sum; s_w[2 * NWARPS + warp] = nkeys; }
    __syncthreads();
    uint32_t before = 0, all_sum = 0, all_keys = 0;
#pragma unroll
    for (int w = 0; w < NWARPS; w++) {
        if (w < warp) before += s_w[w];
        all_sum += s_w[NWARPS + w];
        all_keys += s_w[2 * NWARPS + w];
    }
    const unsigned long long code0 = ((unsigned long long)b << 16) | ((unsigned long long)half << 15);
    if (all_sum == all_keys) {
        if (nz == 0u) return;   // (uniform over the warp; no barrier follows)
        // (32-bit offsets from the half bucket's first entry: a half holds at most 32,768 entries)
        unsigned long long *const co = codes_out + pos;
        uint32_t *const no = counts_out + pos;
        uint32_t wp = before;
        const unsigned long long hi = code0 | (unsigned long long)(2u * warp * PER_WARP);
        for (uint32_t r = 0; r < PER_WARP / 128; r++) {
            const uint4 v = wb[r * 32 + lane];
            const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
            const uint32_t mine = sp16_nonzero_halves(v.x) + sp16_nonzero_halves(v.y) + sp16_nonzero_halves(v.z) + sp16_nonzero_halves(v.w);
            uint32_t incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += t; }
            uint32_t o = wp + (incl - mine);
            const uint32_t cbase = r * 256 + (uint32_t)lane * 8;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const uint32_t c = (j & 1) ? w4[j >> 1] >> 16 : w4[j >> 1] & 0xFFFFu;
                if (c) {
                    co[o] = hi | (unsigned long long)(cbase + j);
                    no[o] = c;
                    o++;
                }
            }
            wp += __shfl_sync(FULL, incl, 31);
        }
        return;
    }
    // ---- a 16-bit counter carried or wrapped: the half again with u32 bins, 16,384 codes at a time ----
    __syncthreads();
    unsigned long long qpos = pos;
    for (uint32_t quarter = 0; quarter < 2; quarter++) {
        for (uint32_t i = threadIdx.x; i < NWORDS / 4; i += THREADS) b4[i] = make_uint4(0u, 0u, 0u, 0u);
        __syncthreads();
        for (int t = ta; t < tb; t++) {
            const uint32_t n = tile_cnt[(size_t)t * SP_BUCKETS + b];
            const uint16_t *gk = fk + tile_place[(size_t)t * SP_BUCKETS + b];
            for (uint32_t i = threadIdx.x; i < n; i += THREADS) {
                const uint32_t v = gk[i];
                if ((v >> 14) == 2u * half + quarter) atomicAdd(bins + (v & 0x3FFFu), 1u);
            }
        }
        __syncthreads();
        uint32_t qnz = 0;
        for (uint32_t r = 0; r < PER_WARP / 128; r++) {
            const uint4 v = wb[r * 32 + lane];
            qnz += (v.x != 0u) + (v.y != 0u) + (v.z != 0u) + (v.w != 0u);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) qnz += __shfl_xor_sync(FULL, qnz, o);
        if (lane == 0) s_w[warp] = qnz;
        __syncthreads();
        uint32_t qbefore = 0, qall = 0;
#pragma unroll
        for (int w = 0; w < NWARPS; w++) { const uint32_t c = s_w[w]; if (w < warp) qbefore += c; qall += c; }
        unsigned long long wp = qpos + qbefore;
        const unsigned long long hi = code0 | ((unsigned long long)quarter << 14) | (unsigned long long)(warp * PER_WARP);
        if (qnz) {
            for (uint32_t r = 0; r < PER_WARP / 128; r++) {
                const uint4 v = wb[r * 32 + lane];
                const uint32_t c4[4] = {v.x, v.y, v.z, v.w};
                const uint32_t mine = (v.x != 0u) + (v.y != 0u) + (v.z != 0u) + (v.w != 0u);
                uint32_t incl = mine;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += t; }
                unsigned long long o = wp + (incl - mine);
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    if (c4[j]) {
                        codes_out[o] = hi | (unsigned long long)(r * 128 + lane * 4 + j);
                        counts_out[o] = c4[j];
                        o++;
                    }
                }
                wp += __shfl_sync(FULL, incl, 31);
            }
        }
        qpos += qall;
        __syncthreads();
    }
}

// ---- bitonic sorting network, all merges ascending (first step of a merge mirrors, the others are strides): with that
// shape, elements beyond n behave as +infinity without being stored, so any n works in place. ----
template <typename KT>
__device__ __forceinline__ void bitonic_sort(KT *a, uint32_t n, uint32_t npad, int tid, int nthreads) {
    const uint32_t pairs = npad >> 1;
    for (uint32_t lk = 1; (1u << lk) <= npad; lk++) {          // merge blocks of K2 = 2^lk
        const uint32_t lh = lk - 1, hmask = (1u << lh) - 1u;
        for (uint32_t t = (uint32_t)tid; t < pairs; t += (uint32_t)nthreads) {
            const uint32_t blk = t >> lh, off = t & hmask;
            const uint32_t i = (blk << lk) + off, j = (blk << lk) + ((1u << lk) - 1u - off);
            if (j < n) { const KT x = a[i], y = a[j]; if (x > y) { a[i] = y; a[j] = x; } }
        }
        __syncthreads();
        for (int lj = (int)lk - 2; lj >= 0; lj--) {              // strides J = 2^lj
            const uint32_t jmask = (1u << lj) - 1u;
            for (uint32_t t = (uint32_t)tid; t < pairs; t += (uint32_t)nthreads) {
                const uint32_t i = ((t >> lj) << (lj + 1)) + (t & jmask), j = i + (1u << lj);
                if (j < n) { const KT x = a[i], y = a[j]; if (x > y) { a[i] = y; a[j] = x; } }
            }
            __syncthreads();
        }
    }
}

// Pass 4.  item = blockIdx.x = file * S + segment; the segment's SP_BUCKETS / S buckets are sorted one after the other.
// nd [item] = number of distinct codes in the item's key range.
template <typename KT, int THREADS>
__global__ void __launch_bounds__(THREADS)
sparse_sort_kernel(KT *__restrict__ keys, const uint64_t *__restrict__ kbase, const uint32_t *__restrict__ boff, uint32_t S,
                   uint32_t *__restrict__ nd) {
    KF_DYN_SMEM(unsigned long long, sp_smem);
    KT *buf = reinterpret_cast<KT *>(sp_smem);
    constexpr uint32_t CAP = 65536 / sizeof(KT);
    __shared__ uint32_t s_heads[THREADS / 32];
    const uint32_t item = blockIdx.x, f = item / S, seg = item - f * S;
    const uint32_t per = SP_BUCKETS / S, b0 = seg * per;
    const uint32_t *bo = boff + (size_t)f * (SP_BUCKETS + 1);
    KT *fk = keys + kbase[f];
    uint32_t heads = 0;   // per thread
    for (uint32_t b = b0; b < b0 + per; b++) {
        const uint32_t s = bo[b], n = bo[b + 1] - s;
        if (n == 0) continue;
        KT *gk = fk + s;
        if (n <= CAP) {
            uint32_t npad = 32;
            while (npad < n) npad <<= 1;
            for (uint32_t i = threadIdx.x; i < npad; i += THREADS) buf[i] = i < n ? gk[i] : (KT)~(KT)0;
            __syncthreads();
            if (n > 1) bitonic_sort<KT>(buf, npad, npad, (int)threadIdx.x, THREADS);
            for (uint32_t i = threadIdx.x; i < n; i += THREADS) {
                const KT x = buf[i];
                gk[i] = x;
                heads += (i == 0 || buf[i - 1] != x) ? 1u : 0u;
            }
            __syncthreads();
        } else {
            // does not fit: the same network on global memory (one CTA: __syncthreads orders its global accesses)
            uint32_t npad = 1u << (32 - __clz((int)(n - 1)));
            if (npad < n) npad = 0x80000000u;   // (n > 2^31 cannot happen: a file holds fewer than 2^32 bytes per bucket scan)
            bitonic_sort<KT>(gk, n, npad, (int)threadIdx.x, THREADS);
            for (uint32_t i = threadIdx.x; i < n; i += THREADS) heads += (i == 0 || gk[i - 1] != gk[i]) ? 1u : 0u;
            __syncthreads();
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) heads += __shfl_xor_sync(FULL, heads, o);
    if ((threadIdx.x & 31) == 0) s_heads[threadIdx.x >> 5] = heads;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < THREADS / 32; w++) t += s_heads[w];
        nd[item] = t;
    }
}

// Pass 5: one CTA.  ooff [n_items + 1] = exclusive scan of nd (64-bit).
__global__ void __launch_bounds__(1024)
sparse_scan_items_kernel(const uint32_t *__restrict__ nd, unsigned long long *__restrict__ ooff, uint32_t n_items) {
    __shared__ unsigned long long wsum[32];
    __shared__ unsigned long long s_base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (uint32_t i0 = 0; i0 < n_items; i0 += 1024) {
        const uint32_t i = i0 + threadIdx.x;
        const unsigned long long mine = i < n_items ? nd[i] : 0ull;
        unsigned long long inc = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned long long t = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            unsigned long long w = wsum[lane], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const unsigned long long t = __shfl_up_sync(FULL, wi, o); if (lane >= o) wi += t; }
            wsum[lane] = wi - w;
        }
        __syncthreads();
        const unsigned long long base = s_base;
        if (i < n_items) ooff[i] = base + wsum[warp] + inc - mine;
        __syncthreads();
        if (threadIdx.x == 1023) s_base = base + wsum[31] + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) ooff[n_items] = s_base;
}

// Pass 6.  The item's key range is sorted (bucket after bucket, and the bucket number is the code's leading bits): its
// runs of equal codes become (code, count) entries at out_base + ooff[item] ...
template <typename KT, int THREADS>
__global__ void __launch_bounds__(THREADS)
sparse_emit_kernel(const KT *__restrict__ keys, const uint64_t *__restrict__ kbase, const uint32_t *__restrict__ boff, uint32_t S,
                   const unsigned long long *__restrict__ ooff, unsigned long long *__restrict__ codes_out, uint32_t *__restrict__ counts_out) {
    __shared__ uint32_t s_idx[THREADS];
    __shared__ unsigned long long s_key[THREADS];
    __shared__ uint32_t wsum[THREADS / 32];
    const uint32_t item = blockIdx.x, f = item / S, seg = item - f * S;
    const uint32_t per = SP_BUCKETS / S;
    const uint32_t *bo = boff + (size_t)f * (SP_BUCKETS + 1);
    const uint32_t r0 = bo[seg * per], r1 = bo[(seg + 1) * per];
    if (r0 == r1) return;
    const KT *fk = keys + kbase[f];
    unsigned long long cur = ooff[item];      // next output slot
    bool pending = false;                     // an open run: (pend_key, pend_start)
    unsigned long long pend_key = 0;
    uint32_t pend_start = 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t t0 = r0; t0 < r1; t0 += THREADS) {
        const uint32_t i = t0 + threadIdx.x;
        const bool valid = i < r1;
        const KT key = valid ? fk[i] : (KT)0;
        const bool head = valid && (i == r0 || fk[i - 1] != key);
        const unsigned bal = __ballot_sync(FULL, head);
        if (lane == 0) wsum[warp] = (uint32_t)__popc(bal);
        __syncthreads();
        uint32_t before = 0, h = 0;
#pragma unroll
        for (int w = 0; w < THREADS / 32; w++) { const uint32_t c = wsum[w]; if (w < warp) before += c; h += c; }
        if (head) {
            const uint32_t pos = before + (uint32_t)__popc(bal & ((1u << lane) - 1u));
            s_idx[pos] = i;
            s_key[pos] = (unsigned long long)key;
        }
        __syncthreads();
        if (h > 0) {
            const uint32_t first = s_idx[0];
            if (pending && threadIdx.x == 0) { codes_out[cur] = pend_key; counts_out[cur] = first - pend_start; }
            const unsigned long long base = cur + (pending ? 1ull : 0ull);
            for (uint32_t j = threadIdx.x; j + 1 < h; j += THREADS) { codes_out[base + j] = s_key[j]; counts_out[base + j] = s_idx[j + 1] - s_idx[j]; }
            cur = base + (h - 1);
            pending = true;
            pend_key = s_key[h - 1];
            pend_start = s_idx[h - 1];
        }
        __syncthreads();
    }
    if (pending && threadIdx.x == 0) { codes_out[cur] = pend_key; counts_out[cur] = r1 - pend_start; }
}

// The FSW fork's k-mer feature matrix (kf2vec/main.py:147-169): one row per observed canonical k-mer -- its k bases as
// float codes A0 T1 C2 G3 (main.py:118), then count / divisor in fp32 (the reference divides the float32 counts by their
// float32 sum: the caller passes that sum).  One thread per matrix element: coalesced stores.
__global__ void __launch_bounds__(256)
sparse_kmer_matrix_kernel(const unsigned long long *__restrict__ codes, const uint32_t *__restrict__ counts, unsigned long long n, int k,
                          float divisor, float *__restrict__ out) {
    const unsigned long long total = n * (unsigned long long)(k + 1);
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long row = i / (unsigned long long)(k + 1);
        const int col = (int)(i - row * (unsigned long long)(k + 1));
        float v;
        if (col < k) {
            const uint32_t d = (uint32_t)(codes[row] >> (2 * (k - 1 - col))) & 3u;   // vocabulary digit A0 C1 G2 T3
            v = (float)((0x1320u >> (4 * d)) & 0xFu);                                // -> A0 T1 C2 G3: 0, 2, 3, 1
        } else {
#ifdef KF_EMU
            v = (float)counts[row] / divisor;
#else
            v = __fdiv_rn((float)counts[row], divisor);
#endif
        }
        out[i] = v;
    }
}

}  // namespace kf
