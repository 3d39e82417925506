// kf_api.cu -- C ABI of libkfcount.so: device context, tile planning, kernel launches, host<->device
// staging.  See include/kfcount.h for the contract and the reference lines each entry point replaces.
#include "kfcount.h"
#include "kf_kernels.cuh"
#include "kf_sparse.cuh"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <queue>
#include <thread>
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

namespace kf {
void canonical_codes(int k, std::vector<uint32_t> &out);  // kf_host.cpp

namespace {

constexpr int THREADS_SMEM = 512;   // 16 warps; 2 CTAs/SM -> 32 warps/SM, 64 regs/thread
constexpr int CTAS_PER_SM = 2;
constexpr int THREADS_GMEM = 512;
constexpr int THREADS_LG = 512;     // line-grid kernel: 1 CTA/SM (128 KiB pair histogram + per-warp TMA staging)
constexpr uint32_t TILE_CHUNKS = 1024;                 // 512 KiB per tile: 64 chunks per warp
constexpr size_t GMEM_WS_LIMIT = (size_t)12 << 30;     // forward-count workspace cap for k >= 8
constexpr int THREADS_PART = 1024;                     // partitioned kernel (k = 8..10): 1 CTA/SM, 128 KiB histogram, 64 regs/thread
constexpr uint64_t PART_MIN_BYTES = 256 << 10;         // smaller files (chunked-mode windows ...) stay on global REDs
constexpr int part_top_bits(int k) { return k == 8 ? 0 : k == 9 ? 3 : 5; }   // PartGeom: 1 / 3 / 11 partitions

struct Ctx {
    int device = -1;
    int sm_count = 0;
    int sm_all = 0;   // the device's SM count; sm_count = SMs the counting kernels are sized for (kf_set_sm_limit)
    uint32_t smem_base = 0;   // shared-window address of dynamic shared memory (kf_smem_base_probe_kernel)
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    std::vector<cudaEvent_t> ev_sub;   // kf_count_buffers: per sub-batch, region zeroed / copies landed
    cudaEvent_t ev_copy = nullptr, ev_k0 = nullptr, ev_k1 = nullptr;   // ev_k0/ev_k1: the current call's pair of the ring below
    bool ev_valid = false;
    // the counting kernels of the last EV_RING calls, timed on the launching stream (read back after the fact, so a
    // timed loop needs no host synchronisation between its steps)
    static constexpr int EV_RING = 64;
    cudaEvent_t ring0[EV_RING] = {nullptr}, ring1[EV_RING] = {nullptr};
    bool ring_valid[EV_RING] = {false};
    long long n_calls = 0;
    // workspaces (grow-only)
    void *d_fwd = nullptr; size_t fwd_cap = 0;
    Tile *d_tiles = nullptr; size_t tiles_cap = 0;
    int *d_cta_begin = nullptr; size_t cta_cap = 0;
    uint32_t *d_canon[KF_MAX_K + 1] = {nullptr};
    uint32_t *d_rank[KF_MAX_K + 1] = {nullptr};    // k >= 8: column of canonical k-mer m (sorted alphabet), else 0xFFFFFFFF
    uint64_t *d_file_off = nullptr; size_t foff_cap = 0;
    uint64_t *d_file_len = nullptr; size_t flen_cap = 0;
    uint8_t *d_formats = nullptr; size_t fmt_cap = 0;
    uint32_t *d_file_P = nullptr; size_t fP_cap = 0;
    uint32_t *d_width_counts = nullptr;
    uint32_t *d_file_row = nullptr; size_t frow_cap = 0;          // first forward-count row of every file (+ total)
    uint32_t *d_cta_first_rank = nullptr; size_t cfr_cap = 0;    // per line-kernel CTA: how many earlier CTAs hold a piece of its first file
    uint32_t pc_rows = 0;
    // k = 8..10: (file, partition) work items of the partitioned shared-memory kernel
    int *d_file_t0 = nullptr; size_t ft0_cap = 0;
    uint32_t *d_items = nullptr; size_t items_cap = 0;
    unsigned int *d_item_counter = nullptr;   // [2]: pass A (or the only pass), pass B
    int pc_items = 0;        // items of the text pass: (file, 0)
    int pc_items_b = 0;      // items of the stream passes: (file, p >= 1), stored after the first pc_items
    uint2 *d_stream = nullptr; size_t stream_cap = 0;
    unsigned long long *d_fold_tot = nullptr; size_t fold_tot_cap = 0;   // per-file totals of the sliced fold   // decoded pieces of pass A: 8 bytes per 16 bytes of arena
    uint32_t pc_part_mode = 0;
    // FASTQ plan: 128 KiB tiles (32 lane ranges), layout-violation offsets
    Tile *d_fq_tiles = nullptr; size_t fq_tiles_cap = 0;
    int *d_fq_cta_begin = nullptr; size_t fq_cta_cap = 0;
    unsigned long long *d_fq_err = nullptr; size_t fq_err_cap = 0;
    int pc_fq_ntiles = 0, pc_fq_nfiles = 0;
    std::vector<unsigned long long> h_fq_err;     // copied back by fetch_status
    int fq_err_n = 0;                             // files covered by d_fq_err in the last call (0: no FASTQ)
    // end-to-end staging
    uint8_t *d_seq = nullptr; size_t seq_cap = 0;              // kf_count_windows: linearised sequence
    uint64_t *d_win_off = nullptr; size_t woff_cap = 0;
    uint32_t *d_win_len = nullptr; size_t wlen_cap = 0;
    uint8_t *d_arena = nullptr; size_t arena_cap = 0;
    unsigned long long *d_counts = nullptr; size_t counts_cap = 0;
    double *d_freq = nullptr; size_t freq_cap = 0;
    unsigned long long *d_totals = nullptr; size_t totals_cap = 0;
    // kf_files_to_kf: two pinned input slabs and two pinned output blocks (grow-only)
    uint8_t *h_slab[2] = {nullptr, nullptr}; size_t h_slab_cap[2] = {0, 0};
    double *h_freq[2] = {nullptr, nullptr}; size_t h_freq_cap[2] = {0, 0};
    unsigned long long *h_cnt[2] = {nullptr, nullptr}; size_t h_cnt_cap[2] = {0, 0};
    // plan cache
    std::vector<uint64_t> pc_offsets, pc_lens;
    std::vector<uint8_t> pc_formats;
    int pc_k = -1, pc_grid = 0, pc_ntiles = 0;
    uint32_t pc_f0 = 0, pc_f1 = 0;
    // plan tables are staged in pinned memory (two blocks used in turn) and uploaded on the launching stream
    uint8_t *h_stage[2] = {nullptr, nullptr}; size_t h_stage_cap[2] = {0, 0};
    cudaEvent_t ev_stage[2] = {nullptr, nullptr};
    bool stage_pending[2] = {false, false};
    int stage_next = 0;
    cudaEvent_t ev_done = nullptr;          // recorded at the end of every run_files on its stream
    cudaStream_t last_stream = nullptr;
    bool last_stream_valid = false;
    std::string last_err;
    int last_launches = 0;
};

Ctx g;
std::mutex g_mu;
void sparse_free_all();   // kf_sparse_host.inc

#define CK(call)                                                                                  \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess) {                                                                  \
            g.last_err = std::string(#call) + ": " + cudaGetErrorString(e_);                      \
            return KF_ERR_CUDA;                                                                   \
        }                                                                                         \
    } while (0)

template <typename T>
int ensure(T *&ptr, size_t &cap, size_t need_bytes) {
    if (need_bytes <= cap) return KF_OK;
    if (ptr) { CK(cudaDeviceSynchronize()); CK(cudaFree(ptr)); ptr = nullptr; cap = 0; }
    size_t want = need_bytes + need_bytes / 8 + 4096;
    cudaError_t e = cudaMalloc((void **)&ptr, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        e = cudaMalloc((void **)&ptr, need_bytes);
        want = need_bytes;
        if (e != cudaSuccess) { cudaGetLastError(); g.last_err = "cudaMalloc failed"; return KF_ERR_NOMEM; }
    }
    cap = want;
    return KF_OK;
}

template <typename T>
int ensure_pinned(T *&ptr, size_t &cap, size_t need_bytes) {
    if (need_bytes <= cap) return KF_OK;
    if (ptr) { CK(cudaFreeHost(ptr)); ptr = nullptr; cap = 0; }
    if (cudaHostAlloc((void **)&ptr, need_bytes, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        g.last_err = "cudaHostAlloc failed";
        return KF_ERR_NOMEM;
    }
    cap = need_bytes;
    return KF_OK;
}

// Host tables -> device, stream-ordered: packed into a pinned staging block (two used in turn) and copied with
// cudaMemcpyAsync on the launching stream, so the kernels queued behind them on that stream see them.  (A blocking
// cudaMemcpy from pageable memory may return before its DMA has landed and is not ordered against non-blocking streams.)
struct Up { void *dst; const void *src; size_t bytes; };
int upload_tables(const std::vector<Up> &ups, cudaStream_t s) {
    size_t total = 0;
    for (auto &u : ups) total += (u.bytes + 255) & ~(size_t)255;
    const int sl = g.stage_next;
    g.stage_next ^= 1;
    if (g.stage_pending[sl]) { CK(cudaEventSynchronize(g.ev_stage[sl])); g.stage_pending[sl] = false; }
    int rc;
    if (total + 256 > g.h_stage_cap[sl] &&
        (rc = ensure_pinned(g.h_stage[sl], g.h_stage_cap[sl], std::max<size_t>(total + total / 2 + 256, (size_t)1 << 20))) != KF_OK) return rc;
    size_t o = 0;
    for (auto &u : ups) {
        memcpy(g.h_stage[sl] + o, u.src, u.bytes);
        CK(cudaMemcpyAsync(u.dst, g.h_stage[sl] + o, u.bytes, cudaMemcpyHostToDevice, s));
        o += (u.bytes + 255) & ~(size_t)255;
    }
    CK(cudaEventRecord(g.ev_stage[sl], s));
    g.stage_pending[sl] = true;
    return KF_OK;
}

int ensure_canon(int k) {
    if (g.d_canon[k]) return KF_OK;
    std::vector<uint32_t> c;
    canonical_codes(k, c);
    CK(cudaMalloc((void **)&g.d_canon[k], c.size() * sizeof(uint32_t)));
    CK(cudaMemcpy(g.d_canon[k], c.data(), c.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    if (k >= 8) {
        // inverse table for the tiled fold: column of every canonical k-mer, 0xFFFFFFFF for the other strand
        std::vector<uint32_t> rank((size_t)1 << (2 * k), 0xFFFFFFFFu);
        for (size_t i = 0; i < c.size(); i++) rank[c[i]] = (uint32_t)i;
        CK(cudaMalloc((void **)&g.d_rank[k], rank.size() * sizeof(uint32_t)));
        CK(cudaMemcpy(g.d_rank[k], rank.data(), rank.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    }
    return KF_OK;
}

// Cut the FASTA files [f0,f1) into per-CTA contiguous, chunk-balanced tile lists.
void build_plan(const uint64_t *offsets, const uint64_t *lens, const uint8_t *formats, uint32_t f0, uint32_t f1,
                int grid, std::vector<Tile> &tiles, std::vector<int> &cta_begin) {
    uint64_t total = 0;
    for (uint32_t f = f0; f < f1; f++)
        if (formats[f] == '>' && lens[f] > 0) total += (lens[f] + CHUNK - 1) / CHUNK;
    tiles.clear();
    cta_begin.assign((size_t)grid + 1, 0);
    uint64_t done = 0;   // chunks assigned so far
    int cta = 0;
    auto cta_hi = [&](int b) { return (total * (uint64_t)(b + 1)) / (uint64_t)grid; };
    for (uint32_t f = f0; f < f1; f++) {
        if (formats[f] != '>' || lens[f] == 0) continue;
        const uint64_t fc0 = offsets[f] / CHUNK;
        const uint64_t nch = (lens[f] + CHUNK - 1) / CHUNK;
        uint64_t pos = 0;
        while (pos < nch) {
            while (cta < grid - 1 && done >= cta_hi(cta)) { cta++; cta_begin[(size_t)cta] = (int)tiles.size(); }
            uint64_t room = cta_hi(cta) - done;
            if (cta == grid - 1) room = total - done;
            uint64_t take = std::min<uint64_t>(std::min<uint64_t>(nch - pos, room), TILE_CHUNKS);
            if (take == 0) take = 1;
            Tile t;
            t.first_chunk = (uint32_t)(fc0 + pos);
            t.n_chunks = (uint32_t)take;
            t.file = f;
            t.file_chunk0 = (uint32_t)fc0;
            tiles.push_back(t);
            pos += take;
            done += take;
        }
    }
    while (cta < grid) { cta++; cta_begin[(size_t)cta] = (int)tiles.size(); }
}

// Forward-count rows.  k <= 7: a FASTA file gets one row per line-kernel CTA that holds a piece of it (that CTA writes
// its row with plain stores; the generic kernel adds into the first row), every other file one row.  The fold sums a
// file's rows.  Line-kernel CTA b takes the tile ranges of the generic CTAs [b * stride, (b + 1) * stride).
void build_rows(const std::vector<Tile> &tiles, const std::vector<int> &cta_begin, int grid, int stride, uint32_t f1,
                bool per_cta_rows, std::vector<uint32_t> &file_row, std::vector<uint32_t> &cta_first_rank) {
    std::vector<uint32_t> nrows((size_t)f1, 1u);
    const int nlc = (grid + stride - 1) / stride;
    cta_first_rank.assign((size_t)std::max(1, nlc), 0u);
    if (per_cta_rows) {
        std::vector<int> last_cta((size_t)f1, -1);
        std::vector<uint32_t> cnt((size_t)f1, 0u);
        for (int b = 0; b < nlc; b++) {
            const int t0 = cta_begin[(size_t)(b * stride)], t1 = cta_begin[(size_t)std::min(grid, (b + 1) * stride)];
            if (t0 < t1) cta_first_rank[(size_t)b] = cnt[tiles[(size_t)t0].file];
            for (int t = t0; t < t1; t++) {
                const uint32_t f = tiles[(size_t)t].file;
                if (last_cta[f] != b) { cnt[f]++; last_cta[f] = b; }
            }
        }
        for (uint32_t f = 0; f < f1; f++) nrows[f] = std::max(1u, cnt[f]);
    }
    file_row.assign((size_t)f1 + 1, 0u);
    for (uint32_t f = 0; f < f1; f++) file_row[f + 1] = file_row[f] + nrows[f];
}

// FASTQ files [f0,f1): tiles of FQ_TILE_CHUNKS chunks in file order, contiguous tile-balanced CTA ranges (and the tile
// range of every FASTQ file).
constexpr uint32_t FQ_TILE_CHUNKS = 32 * FQ_LANE_BYTES / CHUNK;   // 128 KiB: one lane range per lane
void build_fastq_plan(const uint64_t *offsets, const uint64_t *lens, const uint8_t *formats, uint32_t f0, uint32_t f1,
                      int grid, std::vector<Tile> &tiles, std::vector<int> &cta_begin, std::vector<int> &file_tile_begin) {
    tiles.clear();
    file_tile_begin.assign(1, 0);
    for (uint32_t f = f0; f < f1; f++) {
        if (formats[f] != '@' || lens[f] == 0) continue;
        const uint64_t fc0 = offsets[f] / CHUNK;
        const uint64_t nch = (lens[f] + CHUNK - 1) / CHUNK;
        for (uint64_t pos = 0; pos < nch; pos += FQ_TILE_CHUNKS) {
            Tile t;
            t.first_chunk = (uint32_t)(fc0 + pos);
            t.n_chunks = (uint32_t)std::min<uint64_t>(FQ_TILE_CHUNKS, nch - pos);
            t.file = f;
            t.file_chunk0 = (uint32_t)fc0;
            tiles.push_back(t);
        }
        file_tile_begin.push_back((int)tiles.size());
    }
    cta_begin.assign((size_t)grid + 1, 0);
    for (int b = 0; b <= grid; b++) cta_begin[(size_t)b] = (int)((uint64_t)tiles.size() * (uint64_t)b / (uint64_t)grid);
}

constexpr int THREADS_FQ_PAIRS = 1024;   // k = 7 FASTQ: one CTA per SM (192 KB of histograms), 32 warps against the scattered loads

template <int K>
int launch_fastq(const uint8_t *d_arena, int grid, uint32_t file_base, cudaStream_t s) {
    if constexpr (K == 7) {
        // 7-mers counted as 8-mer pairs (half the shared-memory atomics)
        constexpr size_t smem = (32768 + 16384 + 2 * (THREADS_FQ_PAIRS / 32) + 4) * sizeof(uint32_t);
        auto kern = count_fastq_pairs_kernel<THREADS_FQ_PAIRS>;
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid / CTAS_PER_SM, THREADS_FQ_PAIRS, smem, s>>>(d_arena, g.d_fq_tiles, g.d_fq_cta_begin, CTAS_PER_SM, g.d_file_off, g.d_file_len,
                                                               (unsigned long long *)g.d_fwd, g.d_file_row, g.d_fq_err);
    } else if constexpr (K <= KF_MAX_K_SMEM) {
        constexpr size_t smem = sizeof(uint32_t) << (2 * K);
        auto kern = count_fastq_smem_kernel<K, THREADS_SMEM, CTAS_PER_SM>;
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, THREADS_SMEM, smem, s>>>(d_arena, g.d_fq_tiles, g.d_fq_cta_begin, g.d_file_off, g.d_file_len,
                                              (unsigned long long *)g.d_fwd, g.d_file_row, g.d_fq_err);
    } else {
        count_fastq_gmem_kernel<K, THREADS_GMEM><<<grid, THREADS_GMEM, 0, s>>>(d_arena, g.d_fq_tiles, g.d_fq_cta_begin, g.d_file_off, g.d_file_len,
                                                                              (uint32_t *)g.d_fwd, file_base, g.d_fq_err);
    }
    CK(cudaGetLastError());
    return KF_OK;
}

int launch_fastq_k(int k, const uint8_t *d_arena, int grid, uint32_t file_base, cudaStream_t s) {
    switch (k) {
        case 1: return launch_fastq<1>(d_arena, grid, file_base, s);
        case 2: return launch_fastq<2>(d_arena, grid, file_base, s);
        case 3: return launch_fastq<3>(d_arena, grid, file_base, s);
        case 4: return launch_fastq<4>(d_arena, grid, file_base, s);
        case 5: return launch_fastq<5>(d_arena, grid, file_base, s);
        case 6: return launch_fastq<6>(d_arena, grid, file_base, s);
        case 7: return launch_fastq<7>(d_arena, grid, file_base, s);
        case 8: return launch_fastq<8>(d_arena, grid, file_base, s);
        case 9: return launch_fastq<9>(d_arena, grid, file_base, s);
        case 10: return launch_fastq<10>(d_arena, grid, file_base, s);
        case 11: return launch_fastq<11>(d_arena, grid, file_base, s);
        case 12: return launch_fastq<12>(d_arena, grid, file_base, s);
        default: return KF_ERR_ARG;
    }
}

constexpr uint32_t LG_SMEM_BASE = 0x400;   // where dynamic shared memory begins on sm_100 (checked once: kf_init)

template <int LW, uint32_t BASE, bool VIRT = false>
int launch_linegrid_b(const uint8_t *d_arena, int grid_generic, cudaStream_t s) {
    constexpr int NW = THREADS_LG / 32;
    constexpr int SW = LW == 0 ? LN_MAX_LW : LW;   // LW == 0: all widths and the long-line files in one launch
    const size_t smem = lines_kernel_smem<SW, VIRT || LW == 0>(NW);
    static_assert(lines_kernel_smem<SW, VIRT || LW == 0>(NW) <= 227 * 1024, "shared memory of the line kernel");
    auto kern = count_fasta_lines_kernel<LW, THREADS_LG, BASE, VIRT>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid_generic / CTAS_PER_SM, THREADS_LG, smem, s>>>(d_arena, g.d_tiles, g.d_cta_begin, g.d_file_P, g.d_file_off, g.d_file_len,
                                                             (unsigned long long *)g.d_fwd, g.d_file_row, g.d_cta_first_rank, CTAS_PER_SM,
                                                             g.d_width_counts);
    CK(cudaGetLastError());
    return KF_OK;
}
// The pair histogram's REDs carry its shared-window address as an immediate when it is where sm_100 puts it (kf_init
// asked the device); otherwise the variant that adds the base per RED runs.
template <int LW, bool VIRT = false>
int launch_linegrid(const uint8_t *d_arena, int grid_generic, cudaStream_t s) {
    return g.smem_base == LG_SMEM_BASE ? launch_linegrid_b<LW, LG_SMEM_BASE, VIRT>(d_arena, grid_generic, s)
                                       : launch_linegrid_b<LW, 0u, VIRT>(d_arena, grid_generic, s);
}

template <int K>
int launch_smem(const uint8_t *d_arena, int grid, bool force_walker, cudaStream_t s) {
    constexpr size_t smem = sizeof(uint32_t) << (2 * K);
    unsigned long long *fwd = (unsigned long long *)g.d_fwd;
    if (force_walker) {
        auto kern = count_fasta_smem_kernel<K, THREADS_SMEM, CTAS_PER_SM, true>;
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, THREADS_SMEM, smem, s>>>(d_arena, g.d_tiles, g.d_cta_begin, fwd, g.d_file_row, g.d_file_P, g.d_width_counts);
    } else {
        auto kern = count_fasta_smem_kernel<K, THREADS_SMEM, CTAS_PER_SM, false>;
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, THREADS_SMEM, smem, s>>>(d_arena, g.d_tiles, g.d_cta_begin, fwd, g.d_file_row, g.d_file_P, g.d_width_counts);
    }
    CK(cudaGetLastError());
    return KF_OK;
}

template <int K>
int launch_gmem(const uint8_t *d_arena, int grid, bool force_walker, uint32_t file_base, cudaStream_t s) {
    uint32_t *fwd = (uint32_t *)g.d_fwd;
    if constexpr (K >= 8 && K <= 10) {
        if (g.pc_items > 0 && !force_walker) {
            constexpr int PB = part_top_bits(K);
            constexpr size_t smem = PartSink<K, PB>::NWORDS * sizeof(uint32_t);
            CK(cudaMemsetAsync(g.d_item_counter, 0, 2 * sizeof(unsigned int), s));
            if constexpr (PB == 0) {
                auto kern = count_fasta_part_kernel<K, PB, THREADS_PART, 0>;
                CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                kern<<<std::min(g.sm_count, g.pc_items), THREADS_PART, smem, s>>>(d_arena, g.d_tiles, g.d_file_t0, g.d_items, g.pc_items, fwd,
                                                                                 file_base, g.d_item_counter, nullptr);
                CK(cudaGetLastError());
                g.last_launches++;
            } else {
                // pass A parses the text once (partition 0 + the decoded stream), pass B counts the other partitions from it
                auto ka = count_fasta_part_kernel<K, PB, THREADS_PART, 1>;
                auto kb = count_fasta_part_kernel<K, PB, THREADS_PART, 2>;
                CK(cudaFuncSetAttribute(ka, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                CK(cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                ka<<<std::min(g.sm_count, g.pc_items), THREADS_PART, smem, s>>>(d_arena, g.d_tiles, g.d_file_t0, g.d_items, g.pc_items, fwd, file_base,
                                                                               g.d_item_counter, g.d_stream);
                CK(cudaGetLastError());
                kb<<<std::min(g.sm_count, g.pc_items_b), THREADS_PART, smem, s>>>(d_arena, g.d_tiles, g.d_file_t0, g.d_items + g.pc_items, g.pc_items_b, fwd,
                                                                                 file_base, g.d_item_counter + 1, g.d_stream);
                CK(cudaGetLastError());
                g.last_launches += 2;
            }
        }
    }
    const uint32_t *skip = (g.pc_items > 0 && !force_walker) ? g.d_file_P : nullptr;   // files taken by the partitioned kernel
    if (force_walker)
        count_fasta_gmem_kernel<K, THREADS_GMEM, true><<<grid, THREADS_GMEM, 0, s>>>(d_arena, g.d_tiles, g.d_cta_begin, fwd, file_base, skip);
    else
        count_fasta_gmem_kernel<K, THREADS_GMEM, false><<<grid, THREADS_GMEM, 0, s>>>(d_arena, g.d_tiles, g.d_cta_begin, fwd, file_base, skip);
    CK(cudaGetLastError());
    return KF_OK;
}

int launch_count(int k, const uint8_t *d_arena, int grid, bool fw, uint32_t file_base, cudaStream_t s) {
    switch (k) {
        case 1: return launch_smem<1>(d_arena, grid, fw, s);
        case 2: return launch_smem<2>(d_arena, grid, fw, s);
        case 3: return launch_smem<3>(d_arena, grid, fw, s);
        case 4: return launch_smem<4>(d_arena, grid, fw, s);
        case 5: return launch_smem<5>(d_arena, grid, fw, s);
        case 6: return launch_smem<6>(d_arena, grid, fw, s);
        case 7: return launch_smem<7>(d_arena, grid, fw, s);
        case 8: return launch_gmem<8>(d_arena, grid, fw, file_base, s);
        case 9: return launch_gmem<9>(d_arena, grid, fw, file_base, s);
        case 10: return launch_gmem<10>(d_arena, grid, fw, file_base, s);
        case 11: return launch_gmem<11>(d_arena, grid, fw, file_base, s);
        case 12: return launch_gmem<12>(d_arena, grid, fw, file_base, s);
        default: return KF_ERR_ARG;
    }
}

// Count + fold for the files [f0,f1) of a batch whose arena is already on the device.
int run_files(const uint8_t *d_arena, const uint64_t *offsets, const uint64_t *lens, const uint8_t *formats,
              uint32_t f0, uint32_t f1, int k, uint32_t flags, unsigned long long *d_counts, double *d_freq,
              float *d_feat, unsigned long long *d_totals, cudaStream_t s) {
    const bool smem_path = k <= KF_MAX_K_SMEM;
    const size_t NB = (size_t)1 << (2 * k);
    const int64_t V = kf_vocab_size(k);
    const uint32_t nf = f1 - f0;
    const int grid = smem_path ? g.sm_count * CTAS_PER_SM : g.sm_count * 4;
    int rc;
    if ((rc = ensure_canon(k)) != KF_OK) return rc;
    if (smem_path && f0 != 0) return KF_ERR_ARG;   // only the k >= 8 path is run in file batches

    const uint32_t part_mode = (flags & KF_FLAG_NO_LINEGRID) ? 0u : (flags & KF_FLAG_PART_ALL) ? 2u : 1u;
    // plan (cached on layout)
    bool hit = g.pc_k == k && g.pc_grid == grid && g.pc_f0 == f0 && g.pc_f1 == f1 && g.pc_part_mode == part_mode &&
               g.pc_offsets.size() == (size_t)nf && std::equal(g.pc_offsets.begin(), g.pc_offsets.end(), offsets + f0) &&
               std::equal(g.pc_lens.begin(), g.pc_lens.end(), lens + f0) &&
               std::equal(g.pc_formats.begin(), g.pc_formats.end(), formats + f0);
    if (!hit) {
        // Not failure-atomic otherwise: the keys are dropped first and set again only after every upload was queued.
        g.pc_k = -1;
        g.pc_items = 0;
        g.pc_items_b = 0;
        g.pc_part_mode = part_mode;
        // ---- phase 1: host tables ----
        std::vector<Tile> tiles, fq_tiles;
        std::vector<int> cta_begin, fq_cta_begin, fq_ftb, file_t0;
        std::vector<uint32_t> items, taken, file_row, cta_first_rank;
        build_plan(offsets, lens, formats, f0, f1, grid, tiles, cta_begin);
        build_fastq_plan(offsets, lens, formats, f0, f1, grid, fq_tiles, fq_cta_begin, fq_ftb);
        size_t n_items_a = 0, n_items_b = 0;
        if (k >= 8 && k <= 10 && part_mode != 0) {
            // (file, partition) items for FASTA files of at least PART_MIN_BYTES; file_P doubles as the "taken" flag
            file_t0.assign((size_t)f1 + 1, 0);
            {
                size_t t = 0;
                for (uint32_t f = 0; f < f1; f++) {
                    file_t0[f] = (int)t;
                    while (t < tiles.size() && tiles[t].file == f) t++;
                }
                file_t0[f1] = (int)t;
            }
            std::vector<uint32_t> items_b;
            taken.assign((size_t)f1, 0u);
            const uint32_t P = k == 8 ? PartGeom<8, 0>::NPART : k == 9 ? PartGeom<9, 3>::NPART : PartGeom<10, 5>::NPART;
            for (uint32_t f = f0; f < f1; f++) {
                if (formats[f] != '>' || lens[f] == 0 || (lens[f] < PART_MIN_BYTES && part_mode != 2)) continue;
                taken[f] = 1u;
                items.push_back(f << 8);                                               // text pass: partition 0
                for (uint32_t pp = 1; pp < P; pp++) items_b.push_back((f << 8) | pp);  // stream passes, a file's side by side
            }
            if (f1 >= (1u << 24)) items.clear(), items_b.clear(), std::fill(taken.begin(), taken.end(), 0u);   // (file id must fit the item word)
            n_items_a = items.size();
            n_items_b = items_b.size();
            items.insert(items.end(), items_b.begin(), items_b.end());
        }
        if (smem_path) build_rows(tiles, cta_begin, grid, CTAS_PER_SM, f1, k == 7, file_row, cta_first_rank);
        // ---- phase 2: device capacity ----
        if ((rc = ensure(g.d_tiles, g.tiles_cap, (tiles.size() + 1) * sizeof(Tile))) != KF_OK) return rc;
        if ((rc = ensure(g.d_cta_begin, g.cta_cap, cta_begin.size() * sizeof(int))) != KF_OK) return rc;
        // per-file tables are indexed by batch-global file id
        if ((rc = ensure(g.d_file_off, g.foff_cap, (size_t)f1 * sizeof(uint64_t))) != KF_OK) return rc;
        if ((rc = ensure(g.d_file_len, g.flen_cap, (size_t)f1 * sizeof(uint64_t))) != KF_OK) return rc;
        if ((rc = ensure(g.d_formats, g.fmt_cap, (size_t)f1)) != KF_OK) return rc;
        if ((rc = ensure(g.d_file_P, g.fP_cap, (size_t)f1 * sizeof(uint32_t))) != KF_OK) return rc;
        if (!fq_tiles.empty()) {
            if ((rc = ensure(g.d_fq_tiles, g.fq_tiles_cap, fq_tiles.size() * sizeof(Tile))) != KF_OK) return rc;
            if ((rc = ensure(g.d_fq_cta_begin, g.fq_cta_cap, fq_cta_begin.size() * sizeof(int))) != KF_OK) return rc;
        }
        if (!file_t0.empty()) {
            if ((rc = ensure(g.d_file_t0, g.ft0_cap, file_t0.size() * sizeof(int))) != KF_OK) return rc;
            if ((rc = ensure(g.d_items, g.items_cap, (items.size() + 1) * sizeof(uint32_t))) != KF_OK) return rc;
        }
        if (smem_path) {
            if ((rc = ensure(g.d_file_row, g.frow_cap, file_row.size() * sizeof(uint32_t))) != KF_OK) return rc;
            if ((rc = ensure(g.d_cta_first_rank, g.cfr_cap, cta_first_rank.size() * sizeof(uint32_t))) != KF_OK) return rc;
        }
        // ---- phase 3: uploads, stream-ordered.  The tables are packed into a pinned staging block and copied with
        // cudaMemcpyAsync on the launching stream, so the kernels queued behind them on that stream see them (a blocking
        // cudaMemcpy from pageable memory may return before its DMA has landed and is not ordered against non-blocking
        // streams).  A previous call's kernels on ANOTHER stream may still be reading the old tables: wait for them on
        // the device, not on the host. ----
        if (g.last_stream_valid && g.last_stream != s) CK(cudaStreamWaitEvent(s, g.ev_done, 0));
        std::vector<Up> ups;
        auto add = [&](void *dst, const void *src, size_t bytes) { if (bytes) ups.push_back({dst, src, bytes}); };
        add(g.d_tiles, tiles.data(), tiles.size() * sizeof(Tile));
        add(g.d_cta_begin, cta_begin.data(), cta_begin.size() * sizeof(int));
        add(g.d_file_off, offsets, (size_t)f1 * sizeof(uint64_t));
        add(g.d_file_len, lens, (size_t)f1 * sizeof(uint64_t));
        add(g.d_formats, formats, (size_t)f1);
        if (!fq_tiles.empty()) {
            add(g.d_fq_tiles, fq_tiles.data(), fq_tiles.size() * sizeof(Tile));
            add(g.d_fq_cta_begin, fq_cta_begin.data(), fq_cta_begin.size() * sizeof(int));
        }
        if (!file_t0.empty()) {
            add(g.d_file_t0, file_t0.data(), file_t0.size() * sizeof(int));
            add(g.d_items, items.data(), items.size() * sizeof(uint32_t));
            add(g.d_file_P, taken.data(), (size_t)f1 * sizeof(uint32_t));
        }
        if (smem_path) {
            add(g.d_file_row, file_row.data(), file_row.size() * sizeof(uint32_t));
            add(g.d_cta_first_rank, cta_first_rank.data(), cta_first_rank.size() * sizeof(uint32_t));
        }
        if ((rc = upload_tables(ups, s)) != KF_OK) return rc;
        g.pc_fq_ntiles = (int)fq_tiles.size();
        g.pc_fq_nfiles = (int)fq_ftb.size() - 1;
        g.pc_items = (int)n_items_a;
        g.pc_items_b = (int)n_items_b;
        if (smem_path) g.pc_rows = file_row.back();
        g.pc_grid = grid; g.pc_f0 = f0; g.pc_f1 = f1; g.pc_ntiles = (int)tiles.size();
        g.pc_offsets.assign(offsets + f0, offsets + f1);
        g.pc_lens.assign(lens + f0, lens + f1);
        g.pc_formats.assign(formats + f0, formats + f1);
        g.pc_k = k;   // (last: the cache is valid only from here on)
    }
    // forward-count workspace: u64 rows (see build_rows) for k <= 7, one u32 row per file for k >= 8
    const size_t fwd_bytes = smem_path ? (size_t)g.pc_rows * NB * sizeof(unsigned long long) : (size_t)nf * NB * sizeof(uint32_t);
    if ((rc = ensure(g.d_fwd, g.fwd_cap, fwd_bytes)) != KF_OK) return rc;
    if (!smem_path) CK(cudaMemsetAsync(g.d_fwd, 0, fwd_bytes, s));   // (k <= 7: the probe kernel zeroes what needs it)
    if (g.pc_items_b > 0) {
        // decoded stream of the multi-pass path: one uint2 per 16-byte piece up to the batch's last file
        const size_t end = (size_t)((offsets[f1 - 1] + lens[f1 - 1] + CHUNK - 1) / CHUNK + 2);
        if ((rc = ensure(g.d_stream, g.stream_cap, end * 32 * sizeof(uint2))) != KF_OK) return rc;
    }
    const bool force_walker = (flags & KF_FLAG_FORCE_WALKER) != 0;
    const bool use_lg = smem_path && k == 7 && !force_walker && !(flags & KF_FLAG_NO_LINEGRID);
    if (smem_path) {
        if (!g.ev_valid) CK(cudaEventRecord(g.ev_k0, s));
        // line width per file (0 = generic kernel) + zeroing of the rows the line kernel will not write
        CK(cudaMemsetAsync(g.d_width_counts, 0, 8 * sizeof(uint32_t), s));
        probe_line_width_kernel<<<(f1 + 3) / 4, 128, 0, s>>>(d_arena, g.d_file_off, g.d_file_len, g.d_formats, (int)f1,
                                                                use_lg ? 0u : 1u, g.d_file_P, g.d_width_counts,
                                                                (unsigned long long *)g.d_fwd, g.d_file_row, (uint32_t)NB);
        CK(cudaGetLastError());
        CK(cudaEventRecord(g.ev_k1, s));
        g.ev_valid = true;
        g.last_launches++;
    }
    if (g.pc_ntiles > 0) {
        if (!g.ev_valid) CK(cudaEventRecord(g.ev_k0, s));
        if (use_lg) {
            // one launch for every kind of file the line kernel takes: 50 / 60 / 70 / 80 / 100 columns and long lines
            rc = launch_linegrid<0>(d_arena, grid, s);
            if (rc != KF_OK) return rc;
            g.last_launches += 1;
        }
        // file indices inside tiles are batch-global; k >= 8 rows are relative to f0, k <= 7 rows come from d_file_row
        rc = launch_count(k, d_arena, grid, force_walker, f0, s);
        if (rc != KF_OK) return rc;
        CK(cudaEventRecord(g.ev_k1, s));
        g.ev_valid = true;
        g.last_launches++;
    }
    if (g.pc_fq_ntiles > 0) {
        if (!g.ev_valid) CK(cudaEventRecord(g.ev_k0, s));
        if ((rc = launch_fastq_k(k, d_arena, grid, f0, s)) != KF_OK) return rc;
        // files the record-chasing kernel found out of 4-line layout: read front to back, one warp each (exact, slow)
        if (smem_path)
            fastq_multiline_kernel<unsigned long long><<<(nf * 32 + 127) / 128, 128, 0, s>>>(d_arena, g.d_file_off, g.d_file_len, g.d_formats, f0, f1, k,
                                                                                             (unsigned long long *)g.d_fwd, g.d_file_row, g.d_fq_err);
        else
            fastq_multiline_kernel<uint32_t><<<(nf * 32 + 127) / 128, 128, 0, s>>>(d_arena, g.d_file_off, g.d_file_len, g.d_formats, f0, f1, k,
                                                                                   (uint32_t *)g.d_fwd, nullptr, g.d_fq_err);
        CK(cudaGetLastError());
        g.last_launches++;
        CK(cudaEventRecord(g.ev_k1, s));
        g.ev_valid = true;
        g.last_launches += 1;
        g.fq_err_n = (int)f1;
    }
    if (smem_path) {
        // u32 staging is exact when no bin can reach 2^32: a file of fewer than 2^32 bytes holds fewer k-mers than that
        uint64_t max_len = 0;
        for (uint32_t f = f0; f < f1; f++) max_len = std::max(max_len, lens[f]);
        if (max_len < (1ull << 32)) {
            const size_t fsm = NB * sizeof(uint32_t);
            CK(cudaFuncSetAttribute(fold_normalize_smem_kernel<uint32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsm));
            fold_normalize_smem_kernel<uint32_t><<<nf, FOLD_THREADS, fsm, s>>>((const unsigned long long *)g.d_fwd, g.d_canon[k], k, V, flags,
                                                                             use_lg ? g.d_file_P : nullptr, g.d_file_row, d_counts, d_freq, d_feat, d_totals);
        } else {
            const size_t fsm = NB * sizeof(unsigned long long);
            CK(cudaFuncSetAttribute(fold_normalize_smem_kernel<unsigned long long>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsm));
            fold_normalize_smem_kernel<unsigned long long><<<nf, FOLD_THREADS, fsm, s>>>((const unsigned long long *)g.d_fwd, g.d_canon[k], k, V, flags,
                                                                                       use_lg ? g.d_file_P : nullptr, g.d_file_row, d_counts, d_freq, d_feat, d_totals);
        }
    } else if (V > (1 << 14)) {
        // large vocabulary: counts + per-file total first, then the normalisation from the counts just written (the
        // reverse-complement gather is a scattered read: done once); with few files the fold of a file is cut into
        // slices so that every SM has work
        const uint32_t S = std::max<uint32_t>(1u, std::min<uint32_t>(64u, (uint32_t)(2 * g.sm_count) / std::max(1u, nf)));
        if ((rc = ensure(g.d_fold_tot, g.fold_tot_cap, (size_t)nf * sizeof(unsigned long long))) != KF_OK) return rc;
        CK(cudaMemsetAsync(g.d_fold_tot, 0, (size_t)nf * sizeof(unsigned long long), s));
        if (k >= 8 && !(flags & KF_FLAG_NO_LINEGRID))
            fold_counts_tiled_kernel<<<dim3(nf, 1u << (2 * (k - 2 * FOLD_T))), 1024, 0, s>>>((const uint32_t *)g.d_fwd, g.d_rank[k], k, V, f0, d_counts,
                                                                                              g.d_fold_tot);
        else
            fold_counts_sliced_kernel<<<dim3(nf, S), 1024, 0, s>>>((const uint32_t *)g.d_fwd, g.d_canon[k], k, V, f0, d_counts, g.d_fold_tot);
        CK(cudaGetLastError());
        fold_norm_sliced_kernel<<<dim3(nf, S), 1024, 0, s>>>((const uint32_t *)g.d_fwd, g.d_canon[k], k, V, flags, f0, d_counts, d_freq, d_feat,
                                                              d_totals, g.d_fold_tot);
        g.last_launches++;
    } else
        fold_normalize_kernel<uint32_t><<<nf, 1024, 0, s>>>((const uint32_t *)g.d_fwd, g.d_canon[k], k, V, flags, f0, nullptr, nullptr,
                                                             d_counts, d_freq, d_feat, d_totals);
    CK(cudaGetLastError());
    g.last_launches++;
    CK(cudaEventRecord(g.ev_done, s));
    g.last_stream = s;
    g.last_stream_valid = true;
    return KF_OK;
}

int count_device_locked(const uint8_t *d_arena, size_t arena_bytes, const uint64_t *offsets, const uint64_t *lens,
                        const uint8_t *formats, int n, int k, uint32_t flags, unsigned long long *d_counts,
                        double *d_freq, float *d_feat, unsigned long long *d_totals, cudaStream_t s) {
    if (g.device < 0) return KF_ERR_NO_DEVICE;
    if (!d_arena || !offsets || !lens || !formats || n < 0 || k < KF_MIN_K || k > KF_MAX_K) return KF_ERR_ARG;
    uint64_t prev_end = 0;
    for (int i = 0; i < n; i++) {
        if (offsets[i] % CHUNK != 0 || offsets[i] < prev_end) return KF_ERR_LAYOUT;
        prev_end = offsets[i] + lens[i];
    }
    if (prev_end + KF_TAIL_PAD > arena_bytes) return KF_ERR_LAYOUT;
    if ((prev_end + CHUNK - 1) / CHUNK + 2 >= 0xFFFFFFFFull) return KF_ERR_ARG;
    g.last_launches = 0;
    if (g.ev_valid) g.ring_valid[(g.n_calls - 1 + Ctx::EV_RING) % Ctx::EV_RING] = true;   // (the previous call's pair was recorded)
    g.ev_valid = false;
    {
        const int slot = (int)(g.n_calls % Ctx::EV_RING);
        g.n_calls++;
        g.ring_valid[slot] = false;
        g.ev_k0 = g.ring0[slot];
        g.ev_k1 = g.ring1[slot];
    }
    g.fq_err_n = 0;
    if (n == 0) return KF_OK;
    {
        int rc0 = ensure(g.d_fq_err, g.fq_err_cap, (size_t)n * sizeof(unsigned long long));
        if (rc0 != KF_OK) return rc0;
        CK(cudaMemsetAsync(g.d_fq_err, 0xFF, (size_t)n * sizeof(unsigned long long), s));
    }
    const size_t NB = (size_t)1 << (2 * k);
    if (k <= KF_MAX_K_SMEM) return run_files(d_arena, offsets, lens, formats, 0, (uint32_t)n, k, flags, d_counts, d_freq, d_feat, d_totals, s);
    // large k: bound the dense forward-count workspace
    size_t ws_limit = GMEM_WS_LIMIT;
    if (const char *e = getenv("KF_WS_LIMIT_BYTES")) {   // tests: exercise the file batching with a small workspace
        const long long v = atoll(e);
        if (v > 0) ws_limit = (size_t)v;
    }
    size_t per = std::max<size_t>(1, ws_limit / (NB * sizeof(uint32_t)));
    for (uint32_t f0 = 0; f0 < (uint32_t)n; f0 += (uint32_t)per) {
        uint32_t f1 = (uint32_t)std::min<size_t>((size_t)n, (size_t)f0 + per);
        int rc = run_files(d_arena, offsets, lens, formats, f0, f1, k, flags, d_counts, d_freq, d_feat, d_totals, s);
        if (rc != KF_OK) return rc;
    }
    return KF_OK;
}

}  // namespace
}  // namespace kf

using namespace kf;

extern "C" {

int kf_init(int device) {
    std::lock_guard<std::mutex> lk(g_mu);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); g.last_err = "no CUDA device"; return KF_ERR_NO_DEVICE; }
    if (device < 0 || device >= ndev) return KF_ERR_ARG;
    if (g.device == device) return KF_OK;
    if (g.device >= 0) return KF_ERR_ARG;   // one device per process (one process per GPU)
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) { g.last_err = "device is not sm_100 (Blackwell B200)"; return KF_ERR_NO_DEVICE; }
    CK(cudaSetDevice(device));
    CK(cudaStreamCreateWithFlags(&g.stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&g.copy_stream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&g.ev_copy, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&g.ev_done, cudaEventDisableTiming));
    for (int i = 0; i < 2; i++) CK(cudaEventCreateWithFlags(&g.ev_stage[i], cudaEventDisableTiming));
    for (int i = 0; i < Ctx::EV_RING; i++) { CK(cudaEventCreate(&g.ring0[i])); CK(cudaEventCreate(&g.ring1[i])); }
    g.ev_k0 = g.ring0[0];
    g.ev_k1 = g.ring1[0];
    CK(cudaMalloc((void **)&g.d_width_counts, 8 * sizeof(uint32_t)));
    CK(cudaMalloc((void **)&g.d_item_counter, 2 * sizeof(unsigned int)));
    g.sm_count = g.sm_all = prop.multiProcessorCount;
    {
        uint32_t *d_b = nullptr, h_b = 0;
        CK(cudaMalloc((void **)&d_b, sizeof(uint32_t)));
        CK(cudaFuncSetAttribute(kf_smem_base_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        kf_smem_base_probe_kernel<<<1, 512, 200 * 1024, g.stream>>>(d_b);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(&h_b, d_b, sizeof(uint32_t), cudaMemcpyDeviceToHost, g.stream));
        CK(cudaStreamSynchronize(g.stream));
        CK(cudaFree(d_b));
        g.smem_base = h_b;
    }
    g.device = device;
    return KF_OK;
}

int kf_shutdown(void) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g.device < 0) return KF_OK;
    cudaDeviceSynchronize();
    sparse_free_all();
    void *dev_ptrs[] = {g.d_fwd, g.d_tiles, g.d_cta_begin, g.d_arena, g.d_counts, g.d_freq, g.d_totals, g.d_seq, g.d_win_off, g.d_win_len,
                        g.d_fq_tiles, g.d_fq_cta_begin, g.d_fq_err, g.d_file_off, g.d_file_len, g.d_formats, g.d_file_P, g.d_file_row,
                        g.d_cta_first_rank, g.d_width_counts, g.d_file_t0, g.d_items, g.d_item_counter, g.d_stream, g.d_fold_tot};
    for (void *p : dev_ptrs) if (p) cudaFree(p);
    for (auto &p : g.d_canon) { if (p) cudaFree(p); p = nullptr; }
    for (auto &p : g.d_rank) { if (p) cudaFree(p); p = nullptr; }
    for (int i = 0; i < 2; i++) {
        if (g.h_slab[i]) cudaFreeHost(g.h_slab[i]);
        if (g.h_freq[i]) cudaFreeHost(g.h_freq[i]);
        if (g.h_cnt[i]) cudaFreeHost(g.h_cnt[i]);
        if (g.h_stage[i]) cudaFreeHost(g.h_stage[i]);
        if (g.ev_stage[i]) cudaEventDestroy(g.ev_stage[i]);
    }
    for (cudaEvent_t ev : g.ev_sub) cudaEventDestroy(ev);
    cudaStreamDestroy(g.stream); cudaStreamDestroy(g.copy_stream); cudaEventDestroy(g.ev_copy); cudaEventDestroy(g.ev_done);
    for (int i = 0; i < Ctx::EV_RING; i++) { cudaEventDestroy(g.ring0[i]); cudaEventDestroy(g.ring1[i]); }
    cudaGetLastError();
    g = Ctx();
    return KF_OK;
}

int kf_device(void) { return g.device >= 0 ? g.device : KF_ERR_NO_DEVICE; }
const char *kf_last_cuda_error(void) { return g.last_err.c_str(); }
int kf_last_launch_count(void) { return g.last_launches; }
int kf_set_sm_limit(int n_sms) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g.device < 0) return KF_ERR_NO_DEVICE;
    if (n_sms < 0) return KF_ERR_ARG;
    g.sm_count = (n_sms == 0 || n_sms > g.sm_all) ? g.sm_all : n_sms;
    return g.sm_count;
}
int kf_last_count_kernel_ms(float *ms) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g.device < 0) return KF_ERR_NO_DEVICE;
    if (!ms || !g.ev_valid) return KF_ERR_ARG;
    CK(cudaEventSynchronize(g.ev_k1));
    CK(cudaEventElapsedTime(ms, g.ev_k0, g.ev_k1));
    return KF_OK;
}

int kf_count_kernel_ms_history(float *ms_out, int n) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g.device < 0) return KF_ERR_NO_DEVICE;
    if (!ms_out || n < 0) return KF_ERR_ARG;
    if (g.ev_valid) g.ring_valid[(g.n_calls - 1 + Ctx::EV_RING) % Ctx::EV_RING] = true;
    n = std::min<long long>(std::min<long long>(n, Ctx::EV_RING), g.n_calls);
    int w = 0;
    for (int i = n; i >= 1; i--) {   // oldest of the n first
        const int slot = (int)((g.n_calls - i) % Ctx::EV_RING);
        if (!g.ring_valid[slot]) continue;
        CK(cudaEventSynchronize(g.ring1[slot]));
        CK(cudaEventElapsedTime(ms_out + w, g.ring0[slot], g.ring1[slot]));
        w++;
    }
    return w;
}

int kf_count_device(const uint8_t *d_arena, size_t arena_bytes, const uint64_t *offsets, const uint64_t *lens,
                    const uint8_t *formats, int n, int k, uint32_t flags, uint64_t *d_counts, double *d_freq,
                    float *d_feat, uint64_t *d_totals, void *stream) {
    std::lock_guard<std::mutex> lk(g_mu);
    cudaStream_t s = stream ? (cudaStream_t)stream : g.stream;
    return count_device_locked(d_arena, arena_bytes, offsets, lens, formats, n, k, flags,
                               (unsigned long long *)d_counts, d_freq, d_feat, (unsigned long long *)d_totals, s);
}

int kf_count_buffers(const uint8_t *const *bufs, const size_t *lens_in, int n, int k, uint32_t flags,
                     uint64_t *counts_out, double *freq_out, uint64_t *totals_out, int *status_out) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g.device < 0) return KF_ERR_NO_DEVICE;
    if (!bufs || !lens_in || !status_out || n < 0 || k < KF_MIN_K || k > KF_MAX_K) return KF_ERR_ARG;
    if (n == 0) return KF_OK;
    const int64_t V = kf_vocab_size(k);
    std::vector<uint64_t> offsets((size_t)n), lens((size_t)n);
    std::vector<uint8_t> formats((size_t)n);
    uint64_t off = 0;
    for (int i = 0; i < n; i++) {
        offsets[(size_t)i] = off;
        status_out[i] = KF_OK;
        uint64_t L = lens_in[i];
        if (L == 0 || !bufs[i]) { status_out[i] = KF_ERR_EMPTY; L = 0; formats[(size_t)i] = 0; }
        else {
            formats[(size_t)i] = bufs[i][0];
            if (bufs[i][0] != '>' && bufs[i][0] != '@') { status_out[i] = KF_ERR_FORMAT; L = 0; }
        }
        lens[(size_t)i] = L;
        off += (L + CHUNK - 1) / CHUNK * CHUNK;
    }
    const size_t arena_bytes = off + KF_TAIL_PAD;
    int rc;
    if ((rc = ensure(g.d_arena, g.arena_cap, arena_bytes)) != KF_OK) return rc;
    if ((rc = ensure(g.d_counts, g.counts_cap, (size_t)n * V * sizeof(unsigned long long))) != KF_OK) return rc;
    if ((rc = ensure(g.d_freq, g.freq_cap, (size_t)n * V * sizeof(double))) != KF_OK) return rc;
    if ((rc = ensure(g.d_totals, g.totals_cap, (size_t)n * sizeof(unsigned long long))) != KF_OK) return rc;
    // Pipelined over sub-batches of files (about KF_SUB_BATCH_BYTES of text each): the host-to-device copies of
    // sub-batch j + 1 run on the copy stream while sub-batch j is counted and its rows go back -- the call takes the
    // time of its input copies plus the LAST sub-batch's kernels and device-to-host copy, not the sum of the three.
    // A sub-batch's region of the arena is zeroed first (gaps must be NUL) on the counting stream, ahead of the copies.
    static const bool trace = getenv("KF_TRACE_BUFFERS") != nullptr;   // developer: where a call's time goes (stderr)
    const auto tr0 = std::chrono::steady_clock::now();
    auto tr_ms = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tr0).count(); };
    double tr_copies = 0, tr_counts = 0;
    uint64_t sub_bytes = (uint64_t)512 << 20;
    if (const char *e = getenv("KF_SUB_BATCH_BYTES")) { const long long v = atoll(e); if (v > 0) sub_bytes = (uint64_t)v; }
    std::vector<int> sb(1, 0);
    {
        uint64_t acc = 0;
        for (int i = 0; i < n; i++) {
            acc += (lens[(size_t)i] + CHUNK - 1) / CHUNK * CHUNK;
            if (acc >= sub_bytes && i + 1 < n) { sb.push_back(i + 1); acc = 0; }
        }
        sb.push_back(n);
    }
    const int nsb = (int)sb.size() - 1;
    while ((int)g.ev_sub.size() < 2 * nsb) {
        cudaEvent_t ev;
        CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        g.ev_sub.push_back(ev);
    }
    if (g.last_stream_valid && g.last_stream != g.stream) CK(cudaStreamWaitEvent(g.stream, g.ev_done, 0));   // (an earlier call on a caller's stream may still read the arena)
    for (int j = 0; j < nsb; j++) {
        const uint64_t r0 = offsets[(size_t)sb[(size_t)j]];
        const uint64_t r1 = j + 1 < nsb ? offsets[(size_t)sb[(size_t)j + 1]] : arena_bytes;
        CK(cudaMemsetAsync(g.d_arena + r0, 0, r1 - r0, g.stream));
        CK(cudaEventRecord(g.ev_sub[(size_t)(2 * j)], g.stream));
    }
    for (int j = 0; j < nsb; j++) {
        CK(cudaStreamWaitEvent(g.copy_stream, g.ev_sub[(size_t)(2 * j)], 0));
        for (int i = sb[(size_t)j]; i < sb[(size_t)j + 1]; i++)
            if (lens[(size_t)i]) CK(cudaMemcpyAsync(g.d_arena + offsets[(size_t)i], bufs[i], lens[(size_t)i], cudaMemcpyHostToDevice, g.copy_stream));
        CK(cudaEventRecord(g.ev_sub[(size_t)(2 * j + 1)], g.copy_stream));
    }
    tr_copies = tr_ms();
    bool any_fq = false;
    int launches = 0;
    g.h_fq_err.assign((size_t)n, ~0ull);
    for (int j = 0; j < nsb; j++) {
        const int i0 = sb[(size_t)j], ns = sb[(size_t)j + 1] - i0;
        CK(cudaStreamWaitEvent(g.stream, g.ev_sub[(size_t)(2 * j + 1)], 0));
        rc = count_device_locked(g.d_arena, arena_bytes, offsets.data() + i0, lens.data() + i0, formats.data() + i0, ns, k, flags,
                                 g.d_counts + (size_t)i0 * V, g.d_freq + (size_t)i0 * V, nullptr, g.d_totals + i0, g.stream);
        if (rc != KF_OK) { cudaStreamSynchronize(g.copy_stream); cudaStreamSynchronize(g.stream); return rc; }
        launches += g.last_launches;
        if (counts_out) CK(cudaMemcpyAsync(counts_out + (size_t)i0 * V, g.d_counts + (size_t)i0 * V, (size_t)ns * V * sizeof(uint64_t), cudaMemcpyDeviceToHost, g.stream));
        if (freq_out) CK(cudaMemcpyAsync(freq_out + (size_t)i0 * V, g.d_freq + (size_t)i0 * V, (size_t)ns * V * sizeof(double), cudaMemcpyDeviceToHost, g.stream));
        if (totals_out) CK(cudaMemcpyAsync(totals_out + i0, g.d_totals + i0, (size_t)ns * sizeof(uint64_t), cudaMemcpyDeviceToHost, g.stream));
        if (g.fq_err_n > 0) {   // (the next sub-batch reuses d_fq_err: fetch this one's layout reports now, in stream order)
            CK(cudaMemcpyAsync(g.h_fq_err.data() + i0, g.d_fq_err, (size_t)ns * sizeof(unsigned long long), cudaMemcpyDeviceToHost, g.stream));
            any_fq = true;
        }
    }
    tr_counts = tr_ms();
    if (trace) { cudaStreamSynchronize(g.copy_stream); fprintf(stderr, "kf_count_buffers: copies queued %.2f ms, counts queued %.2f ms, copies landed %.2f ms", tr_copies, tr_counts, tr_ms()); }
    CK(cudaStreamSynchronize(g.stream));
    if (trace) fprintf(stderr, ", done %.2f ms (%d sub-batches)\n", tr_ms(), nsb);
    g.last_launches = launches;
    g.fq_err_n = 0;   // (d_fq_err holds the last sub-batch only: nothing for kf_last_file_status to read)
    // 4-line FASTQ layout check: a violation inside the file is an error unless only line ends follow it
    if (any_fq) {
        for (int i = 0; i < n; i++) {
            if (formats[(size_t)i] != '@' || status_out[i] != KF_OK) continue;
            const unsigned long long e = g.h_fq_err[(size_t)i];
            if (e == ~0ull || e < offsets[(size_t)i] || e - offsets[(size_t)i] >= lens[(size_t)i]) continue;
            bool only_eol = true;
            for (uint64_t p = e - offsets[(size_t)i]; p < lens[(size_t)i] && only_eol; p++)
                only_eol = bufs[i][p] == '\n' || bufs[i][p] == '\r';
            if (!only_eol) status_out[i] = KF_ERR_FASTQ;
        }
    }
    return KF_OK;
}

int kf_last_file_status(const uint8_t *d_arena, const uint64_t *offsets, const uint64_t *lens, const uint8_t *formats, int n,
                        int *status_out) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g.device < 0) return KF_ERR_NO_DEVICE;
    if (!status_out || n < 0 || (n > 0 && (!offsets || !lens || !formats))) return KF_ERR_ARG;
    for (int i = 0; i < n; i++)
        status_out[i] = lens[i] == 0 ? KF_ERR_EMPTY : (formats[i] == '>' || formats[i] == '@') ? KF_OK : KF_ERR_FORMAT;
    if (g.fq_err_n <= 0 || n == 0) return KF_OK;
    CK(cudaDeviceSynchronize());
    const int m = std::min(n, g.fq_err_n);
    g.h_fq_err.resize((size_t)m);
    CK(cudaMemcpy(g.h_fq_err.data(), g.d_fq_err, (size_t)m * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    for (int i = 0; i < m; i++) {
        if (formats[i] != '@' || status_out[i] != KF_OK) continue;
        const unsigned long long e = g.h_fq_err[(size_t)i];
        if (e == ~0ull || e < offsets[i] || e - offsets[i] >= lens[i]) continue;
        const uint64_t rest = offsets[i] + lens[i] - e;
        bool only_eol = rest <= 256 && d_arena != nullptr;
        if (only_eol) {
            uint8_t tail[256];
            CK(cudaMemcpy(tail, d_arena + e, (size_t)rest, cudaMemcpyDeviceToHost));
            for (uint64_t p = 0; p < rest && only_eol; p++) only_eol = tail[p] == '\n' || tail[p] == '\r';
        }
        if (!only_eol) status_out[i] = KF_ERR_FASTQ;
    }
    return KF_OK;
}

int kf_count_windows(const uint8_t *seq, size_t seq_len, const uint64_t *win_off, const uint32_t *win_len, int n, int k,
                     uint32_t flags, uint64_t *counts_out, double *freq_out, uint64_t *totals_out) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g.device < 0) return KF_ERR_NO_DEVICE;
    if (n < 0 || k < KF_MIN_K || k > KF_MAX_K || (n > 0 && (!seq || !win_off || !win_len))) return KF_ERR_ARG;
    if (n == 0) return KF_OK;
    uint32_t max_len = 0;
    for (int i = 0; i < n; i++) {
        if (win_off[i] > seq_len || win_len[i] > seq_len - win_off[i]) return KF_ERR_ARG;
        max_len = std::max(max_len, win_len[i]);
    }
    const int64_t V = kf_vocab_size(k);
    const uint64_t slot = ((uint64_t)max_len + 2 + CHUNK - 1) / CHUNK * CHUNK;
    const size_t arena_bytes = (size_t)n * slot + KF_TAIL_PAD;
    int rc;
    if ((rc = ensure(g.d_seq, g.seq_cap, seq_len + 16)) != KF_OK) return rc;
    if ((rc = ensure(g.d_win_off, g.woff_cap, (size_t)n * sizeof(uint64_t))) != KF_OK) return rc;
    if ((rc = ensure(g.d_win_len, g.wlen_cap, (size_t)n * sizeof(uint32_t))) != KF_OK) return rc;
    if ((rc = ensure(g.d_arena, g.arena_cap, arena_bytes)) != KF_OK) return rc;
    if ((rc = ensure(g.d_counts, g.counts_cap, (size_t)n * V * sizeof(unsigned long long))) != KF_OK) return rc;
    if ((rc = ensure(g.d_freq, g.freq_cap, (size_t)n * V * sizeof(double))) != KF_OK) return rc;
    if ((rc = ensure(g.d_totals, g.totals_cap, (size_t)n * sizeof(unsigned long long))) != KF_OK) return rc;
    CK(cudaMemcpyAsync(g.d_seq, seq, seq_len, cudaMemcpyHostToDevice, g.stream));
    CK(cudaMemcpyAsync(g.d_win_off, win_off, (size_t)n * sizeof(uint64_t), cudaMemcpyHostToDevice, g.stream));
    CK(cudaMemcpyAsync(g.d_win_len, win_len, (size_t)n * sizeof(uint32_t), cudaMemcpyHostToDevice, g.stream));
    CK(cudaMemsetAsync(g.d_arena + (size_t)n * slot, 0, KF_TAIL_PAD, g.stream));
    gather_windows_kernel<<<std::min(n, g.sm_count * 8), 256, 0, g.stream>>>(g.d_seq, g.d_win_off, g.d_win_len, n, slot, g.d_arena);
    CK(cudaGetLastError());
    std::vector<uint64_t> offsets((size_t)n), lens((size_t)n);
    std::vector<uint8_t> formats((size_t)n, (uint8_t)'>');
    for (int i = 0; i < n; i++) { offsets[(size_t)i] = (uint64_t)i * slot; lens[(size_t)i] = (uint64_t)win_len[i] + 2; }
    rc = count_device_locked(g.d_arena, arena_bytes, offsets.data(), lens.data(), formats.data(), n, k, flags, g.d_counts,
                             freq_out ? g.d_freq : nullptr, nullptr, g.d_totals, g.stream);
    if (rc != KF_OK) return rc;
    g.last_launches++;
    if (counts_out) CK(cudaMemcpyAsync(counts_out, g.d_counts, (size_t)n * V * sizeof(uint64_t), cudaMemcpyDeviceToHost, g.stream));
    if (freq_out) CK(cudaMemcpyAsync(freq_out, g.d_freq, (size_t)n * V * sizeof(double), cudaMemcpyDeviceToHost, g.stream));
    if (totals_out) CK(cudaMemcpyAsync(totals_out, g.d_totals, (size_t)n * sizeof(uint64_t), cudaMemcpyDeviceToHost, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    return KF_OK;
}

}  // extern "C"

namespace kf {
namespace {

// Minimal worker pool of one kf_files_to_kf call: tasks are file reads and .kf writes.
class Pool {
public:
    explicit Pool(int n) {
        for (int i = 0; i < n; i++) th_.emplace_back([this] { run(); });
    }
    ~Pool() {
        { std::lock_guard<std::mutex> lk(mu_); stop_ = true; }
        cv_.notify_all();
        for (auto &t : th_) t.join();
    }
    void push(std::function<void()> f) {
        { std::lock_guard<std::mutex> lk(mu_); q_.push(std::move(f)); pending_++; }
        cv_.notify_one();
    }
    void wait_all() {
        std::unique_lock<std::mutex> lk(mu_);
        done_cv_.wait(lk, [this] { return pending_ == 0; });
    }
private:
    void run() {
        for (;;) {
            std::function<void()> f;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [this] { return stop_ || !q_.empty(); });
                if (q_.empty()) return;
                f = std::move(q_.front());
                q_.pop();
            }
            f();
            { std::lock_guard<std::mutex> lk(mu_); if (--pending_ == 0) done_cv_.notify_all(); }
        }
    }
    std::vector<std::thread> th_;
    std::queue<std::function<void()>> q_;
    std::mutex mu_;
    std::condition_variable cv_, done_cv_;
    int pending_ = 0;
    bool stop_ = false;
};

double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

}  // namespace
}  // namespace kf

extern "C" {

// Files on disk -> .kf files on disk (the whole of get_frequencies' loop, main.py:301-370), as a three-stage pipeline over
// batches of files: worker threads read batch b+1 into a pinned slab (laid out as the device arena: 512-byte file starts,
// NUL gaps) while the GPU counts batch b (ONE host-to-device copy of the slab, kernels, device-to-host copy of the rows)
// and the workers format and write the .kf rows of batch b-1.
// out_paths / samples may be NULL (no .kf files: only d_feat_out); d_feat_out may be NULL (device float [n][V], the
// trainers' matrix fp32(freq * 1e4), row i for in_paths[i]).
// counts_host / freq_host (kf_count_files): host [n][V] arrays that receive the rows (may be NULL).
static int files_pipeline(const char *const *in_paths, const char *const *out_paths, const char *const *samples, int n, int k,
                          uint32_t flags, int threads, size_t batch_bytes, int *status_out, uint64_t *totals_out, double *stage_seconds,
                          float *d_feat_out, uint64_t *counts_host = nullptr, double *freq_host = nullptr, bool host_rows = false) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g.device < 0) return KF_ERR_NO_DEVICE;
    const bool write_kf = out_paths != nullptr;
    if (!in_paths || (write_kf && !samples) || (!write_kf && !d_feat_out && !host_rows) || !status_out || n < 0 || k < KF_MIN_K || k > KF_MAX_K) return KF_ERR_ARG;
    if (stage_seconds) for (int i = 0; i < 4; i++) stage_seconds[i] = 0.0;
    if (n == 0) return KF_OK;
    const double t_begin = now_s();
    if (threads <= 0) threads = (int)std::max(1u, std::thread::hardware_concurrency());
    threads = std::min(threads, 64);
    if (batch_bytes == 0) batch_bytes = (size_t)256 << 20;
    const int64_t V = kf_vocab_size(k);
    const bool raw = (flags & KF_FLAG_RAW_CNT) != 0, pc = (flags & KF_FLAG_PSEUDOCOUNT) != 0;
    const bool want_counts = write_kf && raw && !pc;   // the reference prints integers only for raw rows with no missing k-mer

    // sizes and batches (file order kept)
    std::vector<uint64_t> fsize((size_t)n, 0);
    for (int i = 0; i < n; i++) {
        struct stat st;
        status_out[i] = KF_OK;
        if (!in_paths[i] || stat(in_paths[i], &st) != 0 || !S_ISREG(st.st_mode)) status_out[i] = KF_ERR_IO;
        else fsize[(size_t)i] = (uint64_t)st.st_size;
        if (totals_out) totals_out[i] = 0;
    }
    struct Batch { int i0, i1; uint64_t bytes; };
    std::vector<Batch> batches;
    {
        int i0 = 0;
        uint64_t cur = 0;
        for (int i = 0; i < n; i++) {
            const uint64_t padded = (fsize[(size_t)i] + CHUNK - 1) / CHUNK * CHUNK;
            if (i > i0 && cur + padded > batch_bytes) { batches.push_back({i0, i, cur}); i0 = i; cur = 0; }
            cur += padded;
        }
        batches.push_back({i0, n, cur});
    }
    const int nb = (int)batches.size();
    std::vector<std::vector<uint64_t>> b_off((size_t)nb), b_len((size_t)nb);
    std::vector<std::vector<uint8_t>> b_fmt((size_t)nb);
    std::vector<std::atomic<int>> io_fail((size_t)n);
    for (auto &a : io_fail) a.store(0);

    kf::Pool pool(threads);
    double t_read = 0, t_gpu = 0, t_write = 0;

    auto issue_reads = [&](int b) -> int {
        const Batch &B = batches[(size_t)b];
        const int sl = b & 1;
        int rc = ensure_pinned(g.h_slab[sl], g.h_slab_cap[sl], (size_t)B.bytes + KF_TAIL_PAD);
        if (rc != KF_OK) return rc;
        auto &off = b_off[(size_t)b];
        auto &len = b_len[(size_t)b];
        off.assign((size_t)(B.i1 - B.i0), 0);
        len.assign((size_t)(B.i1 - B.i0), 0);
        b_fmt[(size_t)b].assign((size_t)(B.i1 - B.i0), 0);
        uint64_t o = 0;
        for (int i = B.i0; i < B.i1; i++) {
            off[(size_t)(i - B.i0)] = o;
            len[(size_t)(i - B.i0)] = status_out[i] == KF_OK ? fsize[(size_t)i] : 0;
            o += (fsize[(size_t)i] + CHUNK - 1) / CHUNK * CHUNK;
        }
        memset(g.h_slab[sl] + B.bytes, 0, KF_TAIL_PAD);
        for (int i = B.i0; i < B.i1; i++) {
            uint8_t *dst = g.h_slab[sl] + off[(size_t)(i - B.i0)];
            const uint64_t L = len[(size_t)(i - B.i0)];
            const uint64_t padded = (fsize[(size_t)i] + CHUNK - 1) / CHUNK * CHUNK;
            const char *path = in_paths[i];
            std::atomic<int> *fail = &io_fail[(size_t)i];
            pool.push([dst, L, padded, path, fail] {
                uint64_t got = 0;
                if (L) {
                    const int fd = open(path, O_RDONLY);
                    if (fd < 0) fail->store(1);
                    else {
                        while (got < L) {
                            const ssize_t r = pread(fd, dst + got, (size_t)std::min<uint64_t>(L - got, (uint64_t)64 << 20), (off_t)got);
                            if (r <= 0) { fail->store(1); break; }
                            got += (uint64_t)r;
                        }
                        close(fd);
                    }
                }
                memset(dst + got, 0, (size_t)(padded - got));   // the gap up to the next file start must read as NUL
            });
        }
        return KF_OK;
    };

    int rc_all = KF_OK;
    {
        const double t0 = now_s();
        rc_all = issue_reads(0);
        pool.wait_all();
        t_read += now_s() - t0;
    }
    for (int b = 0; b < nb && rc_all == KF_OK; b++) {
        const Batch &B = batches[(size_t)b];
        const int sl = b & 1, nf = B.i1 - B.i0;
        auto &off = b_off[(size_t)b];
        auto &len = b_len[(size_t)b];
        auto &fmt = b_fmt[(size_t)b];
        // the reads of this batch are complete (waited for below / above): classify the files
        for (int j = 0; j < nf; j++) {
            const int i = B.i0 + j;
            if (io_fail[(size_t)i].load()) { status_out[i] = KF_ERR_IO; len[(size_t)j] = 0; }
            if (status_out[i] != KF_OK) { fmt[(size_t)j] = 0; len[(size_t)j] = 0; continue; }
            if (len[(size_t)j] == 0) { status_out[i] = KF_ERR_EMPTY; fmt[(size_t)j] = 0; continue; }
            fmt[(size_t)j] = g.h_slab[sl][off[(size_t)j]];
            if (fmt[(size_t)j] != '>' && fmt[(size_t)j] != '@') {
                // rejected: its bytes are already in the slab -- blank its first bytes so that a preceding file that
                // ends without a newline exactly on a 512-byte boundary cannot run on into them
                status_out[i] = KF_ERR_FORMAT;
                memset(g.h_slab[sl] + off[(size_t)j], 0, (size_t)std::min<uint64_t>(64, len[(size_t)j]));
                len[(size_t)j] = 0;
            }
        }
        // next batch's reads and the previous batch's writes run beside this batch's GPU work
        const double t1 = now_s();
        if (b + 1 < nb && (rc_all = issue_reads(b + 1)) != KF_OK) break;
        const size_t arena_bytes = (size_t)B.bytes + KF_TAIL_PAD;
        int rc;
        if ((rc = ensure(g.d_arena, g.arena_cap, arena_bytes)) != KF_OK) { rc_all = rc; break; }
        if ((rc = ensure(g.d_counts, g.counts_cap, (size_t)nf * V * sizeof(unsigned long long))) != KF_OK) { rc_all = rc; break; }
        if ((rc = ensure(g.d_freq, g.freq_cap, (size_t)nf * V * sizeof(double))) != KF_OK) { rc_all = rc; break; }
        if ((rc = ensure(g.d_totals, g.totals_cap, (size_t)nf * sizeof(unsigned long long))) != KF_OK) { rc_all = rc; break; }
        if (write_kf && (rc = ensure_pinned(g.h_freq[sl], g.h_freq_cap[sl], (size_t)nf * V * sizeof(double))) != KF_OK) { rc_all = rc; break; }
        if (want_counts && (rc = ensure_pinned(g.h_cnt[sl], g.h_cnt_cap[sl], (size_t)nf * V * sizeof(unsigned long long))) != KF_OK) { rc_all = rc; break; }
        std::vector<unsigned long long> h_tot((size_t)nf, 0ull);
        CK(cudaMemcpyAsync(g.d_arena, g.h_slab[sl], arena_bytes, cudaMemcpyHostToDevice, g.stream));
        rc = count_device_locked(g.d_arena, arena_bytes, off.data(), len.data(), fmt.data(), nf, k, flags, g.d_counts,
                                 (write_kf || freq_host) ? g.d_freq : nullptr,
                                 d_feat_out ? d_feat_out + (size_t)B.i0 * (size_t)V : nullptr, g.d_totals, g.stream);
        if (rc != KF_OK) { rc_all = rc; break; }
        if (counts_host) CK(cudaMemcpyAsync(counts_host + (size_t)B.i0 * (size_t)V, g.d_counts, (size_t)nf * V * sizeof(uint64_t), cudaMemcpyDeviceToHost, g.stream));
        if (freq_host) CK(cudaMemcpyAsync(freq_host + (size_t)B.i0 * (size_t)V, g.d_freq, (size_t)nf * V * sizeof(double), cudaMemcpyDeviceToHost, g.stream));
        if (write_kf) CK(cudaMemcpyAsync(g.h_freq[sl], g.d_freq, (size_t)nf * V * sizeof(double), cudaMemcpyDeviceToHost, g.stream));
        if (want_counts) CK(cudaMemcpyAsync(g.h_cnt[sl], g.d_counts, (size_t)nf * V * sizeof(unsigned long long), cudaMemcpyDeviceToHost, g.stream));
        CK(cudaMemcpyAsync(h_tot.data(), g.d_totals, (size_t)nf * sizeof(unsigned long long), cudaMemcpyDeviceToHost, g.stream));
        CK(cudaStreamSynchronize(g.stream));
        if (g.fq_err_n > 0) {   // 4-line FASTQ layout check, as in kf_count_buffers
            g.h_fq_err.resize((size_t)nf);
            CK(cudaMemcpy(g.h_fq_err.data(), g.d_fq_err, (size_t)nf * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
            for (int j = 0; j < nf; j++) {
                const int i = B.i0 + j;
                if (fmt[(size_t)j] != '@' || status_out[i] != KF_OK) continue;
                const unsigned long long e = g.h_fq_err[(size_t)j];
                if (e == ~0ull || e < off[(size_t)j] || e - off[(size_t)j] >= len[(size_t)j]) continue;
                bool only_eol = true;
                for (uint64_t p = e; p < off[(size_t)j] + len[(size_t)j] && only_eol; p++)
                    only_eol = g.h_slab[sl][p] == '\n' || g.h_slab[sl][p] == '\r';
                if (!only_eol) status_out[i] = KF_ERR_FASTQ;
            }
        }
        const double t2 = now_s();
        t_gpu += t2 - t1;
        // everything queued before this point (writes of batch b-1, reads of batch b+1) must be done before the output
        // block of batch b-1's parity is reused two batches on; simplest: drain the pool here
        pool.wait_all();
        const double t3 = now_s();
        t_read += t3 - t2;
        for (int j = 0; j < nf; j++) {
            const int i = B.i0 + j;
            if (totals_out) totals_out[i] = h_tot[(size_t)j];
            if (status_out[i] != KF_OK || !write_kf) continue;
            const double *row = g.h_freq[sl] + (size_t)j * V;
            const unsigned long long *crow = want_counts ? g.h_cnt[sl] + (size_t)j * V : nullptr;
            const char *outp = out_paths[i], *smp = samples[i];
            int *st = &status_out[i];
            pool.push([row, crow, outp, smp, st, V] {
                int int_mode = 0;
                if (crow) {
                    int_mode = 1;
                    for (int64_t v = 0; v < V; v++) if (crow[v] == 0) { int_mode = 0; break; }
                }
                const int w = kf_write_kf(outp, smp, row, V, int_mode, 0);
                if (w != KF_OK) *st = w;
            });
        }
    }
    {
        const double t0 = now_s();
        pool.wait_all();
        t_write += now_s() - t0;
    }
    if (stage_seconds) { stage_seconds[0] = t_read; stage_seconds[1] = t_gpu; stage_seconds[2] = t_write; stage_seconds[3] = now_s() - t_begin; }
    return rc_all;
}

int kf_files_to_kf(const char *const *in_paths, const char *const *out_paths, const char *const *samples, int n, int k,
                   uint32_t flags, int threads, size_t batch_bytes, int *status_out, uint64_t *totals_out, double *stage_seconds) {
    if (!out_paths) return KF_ERR_ARG;
    return files_pipeline(in_paths, out_paths, samples, n, k, flags, threads, batch_bytes, status_out, totals_out, stage_seconds, nullptr);
}

// Files on disk -> the [n, V] float32 feature matrix in device memory (fp32(freq * 1e4): the tensor
// train_classifier_model.py:144-150,323 builds from the .kf files), through the same read / GPU pipeline, without the
// text round trip.  Rows of files whose status is not KF_OK are undefined.
// The loop body of main.py:301-342 for n files with the rows returned in host arrays: the same pipelined reads (host
// threads into pinned slabs, one H2D copy per batch) as kf_files_to_kf, no text.
int kf_count_files(const char *const *paths, int n, int k, uint32_t flags, uint64_t *counts_out, double *freq_out,
                   uint64_t *totals_out, int *status_out) {
    return files_pipeline(paths, nullptr, nullptr, n, k, flags, 0, 0, status_out, totals_out, nullptr, nullptr, counts_out, freq_out, true);
}

int kf_files_to_device(const char *const *in_paths, int n, int k, uint32_t flags, int threads, size_t batch_bytes, float *d_feat_out,
                       int *status_out, uint64_t *totals_out, double *stage_seconds) {
    if (!d_feat_out) return KF_ERR_ARG;
    return files_pipeline(in_paths, nullptr, nullptr, n, k, flags, threads, batch_bytes, status_out, totals_out, stage_seconds, d_feat_out);
}

}  // extern "C"

#include "kf_sparse_host.inc"
