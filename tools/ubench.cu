// ubench.cu -- developer micro-benchmarks for the counting kernel's design choices (not product code).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I include -I kf2vecfsw_b200/csrc -I tools tools/ubench.cu kf2vecfsw_b200/csrc/kf_host.cpp tools/kf_synth.cpp -o tools/ubench
// Run on a B200: tools/ubench [arena_MiB]
//
// Every variant streams the same synthetic 80-column FASTA arena (one record) and reports GB/s of file
// bytes, bases per SM-clock (SM clock measured with clock64 inside the kernel) and ms.
#define KF_PIECE_TIMING 1
#include "kf_kernels.cuh"
#include "kfcount.h"
#include "kf_synth.h"
#include <thread>

#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>

using namespace kf;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__global__ void gen_fasta(uint8_t *arena, size_t n, int width) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        uint8_t c;
        if (i < 16) c = (i == 0) ? '>' : (i == 15 ? '\n' : 'x');
        else if ((i - 16) % (size_t)(width + 1) == (size_t)width) c = '\n';
        else {
            uint64_t z = i * 0x9E3779B97F4A7C15ull;
            z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
            z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
            z ^= z >> 31;
            c = "ACGT"[z & 3];
        }
        arena[i] = c;
    }
}

enum Mode { FULL_PROD = 0, XOR_SINK = 1, DECODE_ONLY = 2, LOAD_ONLY = 3, ATOMS_ONLY = 4, PAIR16 = 5, SIMPLE_U32 = 6, ATOMS_ONLY_PAIR = 7 };

// Fast-path-only range processor (clean input assumed) used to compare sinks.
template <int MODE, int K>
__device__ __forceinline__ void simple_range(const uint8_t *__restrict__ arena, uint32_t c0, uint32_t c1, uint32_t *hist,
                                             uint32_t &sink) {
    const int lane = threadIdx.x & 31;
    const uint4 *base = reinterpret_cast<const uint4 *>(arena);
    uint4 wnxt = __ldg(base + (size_t)(c0 + 1) * 32 + lane);
    Lane cur = decode16(__ldg(base + (size_t)c0 * 32 + lane));
    for (uint32_t c = c0; c < c1; ++c) {
        const uint4 wnn = __ldg(base + (size_t)(c + 2) * 32 + lane);
        if (MODE == LOAD_ONLY) { sink ^= wnxt.x ^ wnxt.y ^ wnxt.z ^ wnxt.w; wnxt = wnn; continue; }
        const Lane nxt = decode16(wnxt);
        if (MODE == DECODE_ONLY) { sink ^= cur.bits + cur.n; cur = nxt; wnxt = wnn; continue; }
        const uint32_t xw = cur.bits & ~3u;
        const uint32_t nx0 = nxt.bits & ~3u;
        const uint32_t nb = __shfl_sync(FULL, lane == 0 ? nx0 : xw, (lane + 1) & 31);
        uint32_t hi = cur.bits, lo = nb;
        if (cur.n == 15) { hi |= nb >> 30; lo = nb << 2; }
        if (MODE == PAIR16) {
            // (K+1)-mers at even base offsets; u16 counters packed two per word: word = x >> 1... here
            // word index = low 15 bits, half = top bit (any bijection works; the fold un-permutes)
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
                const uint32_t x = __funnelshift_l(lo, hi, 2 * j) >> (32 - 2 * (K + 1));
                const uint32_t addend = (x >> (2 * (K + 1) - 1)) ? 0x10000u : 1u;
                atomicAdd(hist + (x & ((1u << (2 * (K + 1) - 1)) - 1u)), addend);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 16; j++) {
                if (j < 15 || cur.n == 16) {
                    const uint32_t x = __funnelshift_l(lo, hi, 2 * j) >> (32 - 2 * K);
                    if (MODE == XOR_SINK) sink ^= x * (j + 1);
                    else atomicAdd(hist + x, 1u);
                }
            }
        }
        cur = nxt;
        wnxt = wnn;
    }
}

template <int MODE, int K, int THREADS, int MIN_CTAS, int PF>
__global__ void __launch_bounds__(THREADS, MIN_CTAS)
bench_kernel(const uint8_t *__restrict__ arena, uint32_t n_chunks, unsigned long long *g_out, long long *clk) {
    extern __shared__ uint32_t hist[];
    constexpr int NBINS = (MODE == PAIR16 || MODE == ATOMS_ONLY_PAIR) ? (1 << (2 * (K + 1) - 1)) : (1 << (2 * K));
    for (int i = threadIdx.x; i < NBINS; i += THREADS) hist[i] = 0;
    __syncthreads();
    const long long t0 = clock64();
    const uint32_t nwarps = gridDim.x * (THREADS / 32);
    const uint32_t gw = blockIdx.x * (THREADS / 32) + (threadIdx.x >> 5);
    const uint32_t per = (n_chunks + nwarps - 1) / nwarps;
    const uint32_t c0 = gw * per, c1 = min(c0 + per, n_chunks);
    uint32_t sink = 0;
    if (MODE == ATOMS_ONLY || MODE == ATOMS_ONLY_PAIR) {
        // same number of atomics as the real kernel would issue for this range, random bins, no loads
        uint32_t s = (blockIdx.x * THREADS + threadIdx.x) * 2654435761u + 12345u;
        const int per_iter = (MODE == ATOMS_ONLY) ? 16 : 8;
        for (uint32_t c = c0; c < c1; ++c) {
#pragma unroll
            for (int j = 0; j < per_iter; j++) {
                s = s * 1664525u + 1013904223u;
                const uint32_t x = s >> (32 - ((MODE == ATOMS_ONLY) ? 2 * K : 2 * (K + 1) - 1));
                atomicAdd(hist + x, (MODE == ATOMS_ONLY) ? 1u : ((s & 4096u) ? 0x10000u : 1u));
            }
        }
    } else if (MODE == FULL_PROD) {
        const SmemSink sink = make_smem_sink(hist);
        if (c0 < c1) fasta_process_range<K, false, PF>(GlobalSrc{arena}, c0, c1, 0u, sink);
    } else {
        if (c0 < c1) simple_range<MODE, K>(arena, c0, c1, hist, sink);
    }
    __syncthreads();
    const long long t1 = clock64();
    unsigned long long acc = sink;
    for (int i = threadIdx.x; i < NBINS; i += THREADS) acc += hist[i];
    if (acc == 0x123456789ull) g_out[1] = acc;   // keep everything alive
    atomicAdd(g_out, acc & 0xFFFF);
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int MODE, int K, int THREADS, int MIN_CTAS, int PF = 3>
void run(const char *name, const uint8_t *arena, size_t bytes, int sms, unsigned long long *g_out, long long *d_clk) {
    constexpr size_t smem = sizeof(uint32_t) * ((MODE == PAIR16 || MODE == ATOMS_ONLY_PAIR) ? (1u << (2 * (K + 1) - 1)) : (1u << (2 * K)));
    auto kern = bench_kernel<MODE, K, THREADS, MIN_CTAS, PF>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, smem));
    const int ctas = occ < MIN_CTAS ? occ : MIN_CTAS;
    if (ctas < 1) { printf("%-34s cannot launch (occupancy 0)\n", name); return; }
    const int grid = sms * ctas;
    const uint32_t n_chunks = (uint32_t)(bytes / CHUNK);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    std::vector<long long> clk(grid);
    long long maxclk = 0;
    for (int it = 0; it < 5; it++) {
        CK(cudaEventRecord(e0));
        kern<<<grid, THREADS, smem>>>(arena, n_chunks, g_out, d_clk);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (it >= 1 && ms < best) {
            best = ms;
            CK(cudaMemcpy(clk.data(), d_clk, grid * sizeof(long long), cudaMemcpyDeviceToHost));
            maxclk = 0;
            for (auto c : clk) if (c > maxclk) maxclk = c;
        }
    }
    const double bases = (double)bytes * 80.0 / 81.0;
    const double gbs = (double)bytes / best / 1e6;
    const double bpc = bases / ((double)maxclk * sms);
    printf("%-34s PF=%d thr=%4d ctas/SM=%d  %8.3f ms  %8.1f GB/s  %6.3f Tbases/s  %6.2f bases/clk/SM  clk=%.0f MHz  (%.1f%% of 6550)\n", name, PF, THREADS,
           ctas, best, gbs, bases / best / 1e9, bpc, (double)maxclk / best / 1e3, 100.0 * gbs * 1.02 / 6550.0);
}

// line-grid kernel on the same arena (one file, tiles of 1024 chunks, one CTA per SM)
template <int THREADS>
void run_lg(const uint8_t *arena, size_t bytes, int sms) {
    using G = LineGeom<80>;
    constexpr int NW = THREADS / 32;
    const size_t smem = lines_kernel_smem<80>(NW);
    auto kern = count_fasta_lines_kernel<80, THREADS, 0x400u>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const uint32_t n_chunks = (uint32_t)(bytes / CHUNK);
    std::vector<Tile> tiles; std::vector<int> cta_begin(sms + 1, 0);
    for (int b = 0; b < sms; b++) {
        cta_begin[b] = (int)tiles.size();
        uint32_t lo = (uint32_t)((uint64_t)n_chunks * b / sms), hi = (uint32_t)((uint64_t)n_chunks * (b + 1) / sms);
        for (uint32_t c = lo; c < hi; c += 1024) tiles.push_back(Tile{c, std::min<uint32_t>(1024, hi - c), 0u, 0u});
    }
    cta_begin[sms] = (int)tiles.size();
    Tile *d_tiles; int *d_cb; uint32_t *d_P; uint64_t *d_off, *d_len; unsigned long long *d_fwd, *d_scr;
    CK(cudaMalloc(&d_tiles, tiles.size() * sizeof(Tile))); CK(cudaMemcpy(d_tiles, tiles.data(), tiles.size() * sizeof(Tile), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_cb, cta_begin.size() * sizeof(int))); CK(cudaMemcpy(d_cb, cta_begin.data(), cta_begin.size() * sizeof(int), cudaMemcpyHostToDevice));
    uint32_t P = 81; uint64_t off = 0, len = bytes;
    CK(cudaMalloc(&d_P, 4)); CK(cudaMemcpy(d_P, &P, 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_off, 8)); CK(cudaMemcpy(d_off, &off, 8, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_len, 8)); CK(cudaMemcpy(d_len, &len, 8, cudaMemcpyHostToDevice));
    uint32_t *d_wc; uint32_t wc[4] = {1, 1, 1, 1}; CK(cudaMalloc(&d_wc, 16)); CK(cudaMemcpy(d_wc, wc, 16, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_fwd, (size_t)sms * 16384 * 8)); CK(cudaMemset(d_fwd, 0, (size_t)sms * 16384 * 8));
    CK(cudaMalloc(&d_scr, (size_t)sms * 16384 * 8)); CK(cudaMemset(d_scr, 0, (size_t)sms * 16384 * 8));
    uint32_t *d_row, *d_fc; uint32_t rows[2] = {0u, (uint32_t)sms};
    std::vector<uint32_t> ranks(sms); for (int b = 0; b < sms; b++) ranks[b] = b;
    CK(cudaMalloc(&d_row, 8)); CK(cudaMemcpy(d_row, rows, 8, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_fc, 4 * sms)); CK(cudaMemcpy(d_fc, ranks.data(), 4 * sms, cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int it = 0; it < 5; it++) {
        CK(cudaMemset(d_fwd, 0, (size_t)sms * 16384 * 8));
        CK(cudaEventRecord(e0));
        kern<<<sms, THREADS, smem>>>(arena, d_tiles, d_cb, d_P, d_off, d_len, d_fwd, d_row, d_fc, 1, d_wc);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (it >= 1 && ms < best) best = ms;
    }
    std::vector<unsigned long long> h((size_t)sms * 16384);
    CK(cudaMemcpy(h.data(), d_fwd, (size_t)sms * 16384 * 8, cudaMemcpyDeviceToHost));
    unsigned long long tot = 0; for (auto v : h) tot += v;
    const double bases = (double)bytes * 80.0 / 81.0, gbs = (double)bytes / best / 1e6;
    printf("%-34s      thr=%4d ctas/SM=1  %8.3f ms  %8.1f GB/s  %6.3f Tbases/s  %6.2f bases/clk/SM@1.965GHz  total 7-mers %llu  (%.1f%% of 6550)\n",
           "line kernel LW=80 pair16", THREADS, best, gbs, bases / best / 1e9, bases / (best * 1e-3 * 1.965e9 * sms), tot, 100.0 * gbs * 1.02 / 6550.0);
    cudaFree(d_tiles); cudaFree(d_cb); cudaFree(d_P); cudaFree(d_off); cudaFree(d_len); cudaFree(d_fwd); cudaFree(d_scr);
}

// many equal files (16-byte header, 80-column lines, short last line) through the line kernel, with the per-piece
// cycle breakdown of KF_PIECE_TIMING
__global__ void gen_fasta_files(uint8_t *arena, size_t file_bytes, size_t slot, int nfiles) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = slot * (size_t)nfiles, stride = (size_t)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        const size_t o = i % slot;
        uint8_t c = 0;
        if (o < file_bytes) {
            if (o < 16) c = (o == 0) ? '>' : (o == 15 ? '\n' : 'x');
            else if ((o - 16) % 81 == 80 || o == file_bytes - 1) c = '\n';
            else {
                uint64_t z = i * 0x9E3779B97F4A7C15ull;
                z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
                z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
                z ^= z >> 31;
                c = "ACGT"[z & 3];
            }
        }
        arena[i] = c;
    }
}

template <int THREADS>
void run_lg_files(uint8_t *arena, size_t arena_bytes, int nfiles, int sms, int max_contigs = 0, int n_runs = 0, long long bases = 0) {
    using G = LineGeom<80>;
    constexpr int NW = THREADS / 32;
    const size_t smem = lines_kernel_smem<80>(NW);
    auto kern = count_fasta_lines_kernel<80, THREADS, 0x400u>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    size_t slot = (arena_bytes / nfiles) / CHUNK * CHUNK;
    size_t file_bytes = slot - 100;
    std::vector<uint64_t> flen;
    if (max_contigs > 0) {
        // realistic files from the library's generator (contigs, N runs); every file in a slot of the largest size
        flen.resize(nfiles);
        size_t mx = 0;
        for (int f = 0; f < nfiles; f++) { flen[f] = (uint64_t)kf_synth_fasta_ex(1, f, bases, 80, max_contigs, n_runs, nullptr, 0); mx = std::max<size_t>(mx, flen[f]); }
        slot = (mx + CHUNK - 1) / CHUNK * CHUNK;
        if ((size_t)nfiles * slot > arena_bytes) { printf("arena too small for synthetic mode\n"); return; }
        std::vector<uint8_t> host((size_t)nfiles * slot, 0);
        std::vector<std::thread> th;
        for (int w = 0; w < 16; w++) th.emplace_back([&, w]() { for (int f = w; f < nfiles; f += 16) kf_synth_fasta_ex(1, f, bases, 80, max_contigs, n_runs, host.data() + (size_t)f * slot, flen[f]); });
        for (auto &t : th) t.join();
        CK(cudaMemcpy(arena, host.data(), host.size(), cudaMemcpyHostToDevice));
        file_bytes = mx;
    } else {
        gen_fasta_files<<<sms * 8, 256>>>(arena, file_bytes, slot, nfiles);
    }
    CK(cudaDeviceSynchronize());
    const uint64_t total_chunks = (uint64_t)nfiles * (slot / CHUNK);
    std::vector<Tile> tiles; std::vector<int> cta_begin(sms + 1, 0);
    std::vector<uint64_t> off(nfiles), len(nfiles); std::vector<uint32_t> P(nfiles, 81);
    for (int f = 0; f < nfiles; f++) { off[f] = (uint64_t)f * slot; len[f] = flen.empty() ? file_bytes : flen[f]; }
    {
        uint64_t done = 0; int cta = 0;
        auto hi = [&](int b) { return total_chunks * (uint64_t)(b + 1) / (uint64_t)sms; };
        for (int f = 0; f < nfiles; f++) {
            uint64_t fc0 = off[f] / CHUNK, nch = slot / CHUNK, pos = 0;
            while (pos < nch) {
                while (cta < sms - 1 && done >= hi(cta)) { cta++; cta_begin[cta] = (int)tiles.size(); }
                uint64_t room = (cta == sms - 1) ? total_chunks - done : hi(cta) - done;
                uint64_t take = std::min<uint64_t>(std::min<uint64_t>(nch - pos, room), 1024);
                if (!take) take = 1;
                tiles.push_back(Tile{(uint32_t)(fc0 + pos), (uint32_t)take, (uint32_t)f, (uint32_t)fc0});
                pos += take; done += take;
            }
        }
        while (cta < sms) { cta++; cta_begin[cta] = (int)tiles.size(); }
    }
    Tile *d_tiles; int *d_cb; uint32_t *d_P; uint64_t *d_off, *d_len; unsigned long long *d_fwd, *d_scr;
    CK(cudaMalloc(&d_tiles, tiles.size() * sizeof(Tile))); CK(cudaMemcpy(d_tiles, tiles.data(), tiles.size() * sizeof(Tile), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_cb, cta_begin.size() * sizeof(int))); CK(cudaMemcpy(d_cb, cta_begin.data(), cta_begin.size() * sizeof(int), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_P, 4 * nfiles)); CK(cudaMemcpy(d_P, P.data(), 4 * nfiles, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_off, 8 * nfiles)); CK(cudaMemcpy(d_off, off.data(), 8 * nfiles, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_len, 8 * nfiles)); CK(cudaMemcpy(d_len, len.data(), 8 * nfiles, cudaMemcpyHostToDevice));
    uint32_t *d_wc; uint32_t wc[4] = {1, 1, 1, 1}; CK(cudaMalloc(&d_wc, 16)); CK(cudaMemcpy(d_wc, wc, 16, cudaMemcpyHostToDevice));
    std::vector<uint32_t> frow(nfiles + 1, 0), ffc(sms, 0);
    {
        std::vector<uint32_t> cnt(nfiles, 0); std::vector<int> last(nfiles, -1);
        for (int b = 0; b < sms; b++) {
            if (cta_begin[b] < cta_begin[b + 1]) ffc[b] = cnt[tiles[cta_begin[b]].file];
            for (int t = cta_begin[b]; t < cta_begin[b + 1]; t++) {
                uint32_t f = tiles[t].file;
                if (last[f] != b) { cnt[f]++; last[f] = b; }
            }
        }
        for (int f = 0; f < nfiles; f++) frow[f + 1] = frow[f] + std::max(1u, cnt[f]);
    }
    const size_t nrows = frow[nfiles];
    uint32_t *d_row, *d_fc;
    CK(cudaMalloc(&d_row, 4 * (nfiles + 1))); CK(cudaMemcpy(d_row, frow.data(), 4 * (nfiles + 1), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_fc, 4 * sms)); CK(cudaMemcpy(d_fc, ffc.data(), 4 * sms, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_fwd, nrows * 16384 * 8));
    CK(cudaMalloc(&d_scr, (size_t)sms * 16384 * 8)); CK(cudaMemset(d_scr, 0, (size_t)sms * 16384 * 8));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    unsigned long long tm[16] = {0}, zero[16] = {0};
    for (int it = 0; it < 4; it++) {
        CK(cudaMemset(d_fwd, 0, nrows * 16384 * 8));
        CK(cudaMemcpyToSymbol(g_piece_timing, zero, sizeof zero));
        CK(cudaEventRecord(e0));
        kern<<<sms, THREADS, smem>>>(arena, d_tiles, d_cb, d_P, d_off, d_len, d_fwd, d_row, d_fc, 1, d_wc);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (it >= 1 && ms < best) { best = ms; CK(cudaMemcpyFromSymbol(tm, g_piece_timing, sizeof tm)); }
    }
    const double bytes = (double)nfiles * file_bytes, pieces = (double)tm[3];
    printf("files=%5d x %8zu B  %7.3f ms  %7.1f GB/s | per CTA: total %.0f kclk, pieces %.1f | per piece: units %.1f kclk/warp, wait+checksum %.1f kclk/warp, "
           "fold %.1f kclk/warp, zero %.1f kclk/warp, anchor %.1f kclk\n", nfiles, file_bytes, best, bytes / best / 1e6, tm[4] / (double)sms / 1e3, pieces / sms,
           tm[0] / pieces / NW / 1e3, tm[1] / pieces / NW / 1e3, tm[6] / pieces / NW / 1e3, tm[2] / pieces / NW / 1e3, tm[5] / pieces / 1e3);
    printf("      windows: clean %llu x %.2f kclk | with a dirty lane %llu x %.2f kclk | grid breaks %llu: generic region %.2f kclk + next-line search %.2f kclk each\n",
           tm[8], tm[8] ? tm[9] / (double)tm[8] / 1e3 : 0.0, tm[10], tm[10] ? tm[11] / (double)tm[10] / 1e3 : 0.0, tm[12],
           tm[12] ? tm[13] / (double)tm[12] / 1e3 : 0.0, tm[12] ? tm[14] / (double)tm[12] / 1e3 : 0.0);
    cudaFree(d_tiles); cudaFree(d_cb); cudaFree(d_P); cudaFree(d_off); cudaFree(d_len); cudaFree(d_fwd); cudaFree(d_scr); cudaFree(d_wc);
}

int main(int argc, char **argv) {
    size_t mib = argc > 1 ? (size_t)atol(argv[1]) : 1024;
    size_t bytes = mib << 20;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    printf("device %s  SMs %d  max clk %d MHz  arena %zu MiB\n", prop.name, prop.multiProcessorCount, prop.clockRate / 1000, mib);
    uint8_t *arena;
    CK(cudaMalloc(&arena, bytes + 4 * CHUNK));
    CK(cudaMemset(arena, 0, bytes + 4 * CHUNK));
    gen_fasta<<<prop.multiProcessorCount * 8, 256>>>(arena, bytes, 80);
    CK(cudaDeviceSynchronize());
    unsigned long long *g_out; long long *d_clk;
    CK(cudaMalloc(&g_out, 16)); CK(cudaMemset(g_out, 0, 16));
    CK(cudaMalloc(&d_clk, sizeof(long long) * 4096));
    const int sms = prop.multiProcessorCount;
    run<LOAD_ONLY, 7, 512, 2>("load only (LDG.128 stream)", arena, bytes, sms, g_out, d_clk);
    run<DECODE_ONLY, 7, 512, 2>("load+decode16", arena, bytes, sms, g_out, d_clk);
    run<XOR_SINK, 7, 512, 2>("load+decode+extract (xor sink)", arena, bytes, sms, g_out, d_clk);
    run<ATOMS_ONLY, 7, 512, 2>("atomics only u32 16384 bins", arena, bytes, sms, g_out, d_clk);
    run<SIMPLE_U32, 7, 512, 2>("simple fast path u32", arena, bytes, sms, g_out, d_clk);
    run<FULL_PROD, 7, 512, 2, 3>("production range processor", arena, bytes, sms, g_out, d_clk);
    run_lg<512>(arena, bytes, sms);
    for (int nf : {148, 1000, 4000}) run_lg_files<512>(arena, bytes, nf, sms);
    run_lg_files<512>(arena, bytes, 296, sms, 1, 0, 5000000);
    run_lg_files<512>(arena, bytes, 296, sms, 50, 0, 5000000);
    run_lg_files<512>(arena, bytes, 296, sms, 1, 10, 5000000);
    run_lg_files<512>(arena, bytes, 296, sms, 50, 10, 5000000);
    printf("done\n");
    return 0;
}
