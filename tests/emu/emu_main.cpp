// emu_main.cpp -- TEST INFRASTRUCTURE.  Runs the real kernel code of kf_kernels.cuh under the host
// emulation (cuda_emu.h) with the real tile planner's layout rules, on one or more FASTA files, and
// prints the canonical counts (one line per file) so tests/ can compare them with the oracle.
//   usage: emu_main <k> <threads_per_cta> <grid> <force_walker 0|1> <tile_chunks> <use_linegrid 0|1> file...
#include "cuda_emu.h"
#include "../../kf2vecfsw_b200/csrc/kf_kernels.cuh"
#include "../../kf2vecfsw_b200/csrc/kf_sparse.cuh"

#include <cstdio>
#include <cstdlib>
#include <string>

namespace emu {
thread_local dim3 t_threadIdx, t_blockIdx, t_blockDim, t_gridDim;
thread_local BlockShared *t_block = nullptr;
void launch(unsigned grid, unsigned block, size_t smem_bytes, const std::function<void()> &fn) {
    for (unsigned b = 0; b < grid; b++) {
        BlockShared bs;
        pthread_barrier_init(&bs.bar, nullptr, block);
        bs.smem.assign(smem_bytes + 16, 0);
        bs.warps = std::vector<WarpShared>(block / 32);
        for (auto &w : bs.warps) pthread_barrier_init(&w.bar, nullptr, 32);
        std::vector<std::thread> th;
        for (unsigned t = 0; t < block; t++)
            th.emplace_back([&, t]() {
                t_threadIdx.x = t; t_blockIdx.x = b; t_blockDim.x = block; t_gridDim.x = grid;
                t_block = &bs;
                fn();
            });
        for (auto &x : th) x.join();
    }
}
}  // namespace emu

namespace kf { void canonical_codes(int k, std::vector<uint32_t> &out); }
#ifdef KF_EMU_RANDOM_UNITS
namespace kf { unsigned g_emu_seed = 1; }   // unit sizes of the line kernels are drawn from it (env KF_EMU_SEED)
#endif

using namespace kf;

template <int K, int THREADS>
static void run(const uint8_t *arena, const std::vector<Tile> &tiles, const std::vector<int> &cta_begin, int grid,
                bool fw, unsigned long long *fwd, const uint32_t *file_row, const uint32_t *file_P, const uint32_t *wc) {
    size_t smem = sizeof(uint32_t) << (2 * K);
    if (fw) emu::launch(grid, THREADS, smem, [&]() { count_fasta_smem_kernel<K, THREADS, 1, true>(arena, tiles.data(), cta_begin.data(), fwd, file_row, file_P, wc); });
    else    emu::launch(grid, THREADS, smem, [&]() { count_fasta_smem_kernel<K, THREADS, 1, false>(arena, tiles.data(), cta_begin.data(), fwd, file_row, file_P, wc); });
}

template <int LW, int THREADS, bool VIRT = false>
static void run_lg(const uint8_t *arena, const std::vector<Tile> &tiles, const std::vector<int> &cta_begin, int grid,
                   const uint32_t *file_P, const uint64_t *off, const uint64_t *len, unsigned long long *fwd,
                   const uint32_t *file_row, const uint32_t *file_first_cta, const uint32_t *wc) {
    constexpr int NW = THREADS / 32;
    size_t smem = lines_kernel_smem<(LW == 0 ? LN_MAX_LW : LW), (VIRT || LW == 0)>(NW);
    emu::launch(grid, THREADS, smem, [&]() {
        count_fasta_lines_kernel<LW, THREADS, 0u, VIRT>(arena, tiles.data(), cta_begin.data(), file_P, off, len, fwd, file_row, file_first_cta, 1, wc);
    });
}

int main(int argc, char **argv) {
    if (argc < 8) { fprintf(stderr, "usage\n"); return 2; }
#ifdef KF_EMU_RANDOM_UNITS
    if (const char *e = getenv("KF_EMU_SEED")) kf::g_emu_seed = (unsigned)atoi(e);
#endif
    int k = atoi(argv[1]), threads = atoi(argv[2]), grid = atoi(argv[3]);
    bool fw = atoi(argv[4]) != 0;
    uint32_t tile_chunks = (uint32_t)atoi(argv[5]);
    const int mode = atoi(argv[6]);   // 0 generic, 1 line kernel allowed, 2 partitioned kernel (k = 4, 5 stand in for 9, 10)
    bool use_lg = mode == 1;
    argv += 1; argc -= 1;
    int n = argc - 6;
    std::vector<uint64_t> off(n), len(n);
    std::vector<uint8_t> arena;
    for (int i = 0; i < n; i++) {
        FILE *f = fopen(argv[6 + i], "rb");
        if (!f) { perror(argv[6 + i]); return 1; }
        fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET);
        off[i] = arena.size();
        arena.resize(arena.size() + (size_t)(sz + CHUNK - 1) / CHUNK * CHUNK, 0);
        if (fread(arena.data() + off[i], 1, (size_t)sz, f) != (size_t)sz) return 1;
        fclose(f);
        len[i] = (uint64_t)sz;
    }
    arena.resize(arena.size() + 4096, 0);
    // same cutting rule as kf_api.cu:build_plan (chunk-balanced contiguous CTA ranges, tiles <= tile_chunks)
    uint64_t total = 0;
    for (int i = 0; i < n; i++) if (!len[i] || arena[off[i]] == '>') total += (len[i] + CHUNK - 1) / CHUNK;
    std::vector<Tile> tiles; std::vector<int> cta_begin(grid + 1, 0);
    uint64_t done = 0; int cta = 0;
    auto cta_hi = [&](int b) { return total * (uint64_t)(b + 1) / (uint64_t)grid; };
    for (int f = 0; f < n; f++) {
        if (len[f] && arena[off[f]] != '>') continue;   // FASTA plan only (kf_api.cu:build_plan)
        uint64_t fc0 = off[f] / CHUNK, nch = (len[f] + CHUNK - 1) / CHUNK, pos = 0;
        while (pos < nch) {
            while (cta < grid - 1 && done >= cta_hi(cta)) { cta++; cta_begin[cta] = (int)tiles.size(); }
            uint64_t room = (cta == grid - 1) ? total - done : cta_hi(cta) - done;
            uint64_t take = std::min<uint64_t>(std::min<uint64_t>(nch - pos, room), tile_chunks);
            if (!take) take = 1;
            tiles.push_back(Tile{(uint32_t)(fc0 + pos), (uint32_t)take, (uint32_t)f, (uint32_t)fc0});
            pos += take; done += take;
        }
    }
    while (cta < grid) { cta++; cta_begin[cta] = (int)tiles.size(); }
    if (mode == 3) {
        // sparse sort-and-run-length path, the launch sequence of kf_sparse_host.inc:sparse_run_batch (one sub-batch)
        const int thr = threads == 512 ? 64 : threads;
        std::vector<int> file_t0(n + 1, 0);
        { size_t t = 0; for (int f = 0; f < n; f++) { file_t0[f] = (int)t; while (t < tiles.size() && tiles[t].file == (uint32_t)f) t++; } file_t0[n] = (int)t; }
        const int R = 2 * k - SP_BUCKET_BITS;
        const bool hist_path = R <= SP_HIST_MAX_R && k <= 16;
        const bool path16 = hist_path && 2 * k > 16 && 2 * k <= 24 && !getenv("KF_SPARSE_NO16");   // (kf_sparse_host.inc)
        const uint32_t nbk16 = path16 ? 1u << (2 * k - 16) : 0u;
        std::vector<uint64_t> kbase(n + 1, 0);
        for (int f = 0; f < n; f++) {
            kbase[f + 1] = kbase[f] + ((len[f] && arena[off[f]] == '>') ? len[f] : 0);
            if (path16) kbase[f + 1] = (kbase[f + 1] + 8ull * nbk16 * (uint64_t)(file_t0[f + 1] - file_t0[f]) + 7ull) & ~7ull;
        }
        std::vector<uint32_t> tile_place((tiles.size() + 1) * SP_BUCKETS, 0xBEEFu);
        const int bshift = path16 ? 16 : -1;
        uint32_t S = 1;
        while (S < SP_BUCKETS / 8 && (uint64_t)n * S < 16) S <<= 1;
        const uint32_t n_items = hist_path ? (uint32_t)n * SP_BUCKETS : (uint32_t)n * S;
        std::vector<uint32_t> tile_hist((tiles.size() + 1) * SP_BUCKETS, 0xDEADu), boff((size_t)n * (SP_BUCKETS + 1), 0u), nd(n_items, 0u), ooff32(n_items, 0u);
        std::vector<unsigned long long> ooff(n_items + 1, 0ull), totals(n, 0ull), nd_file(n, 0ull), out_base(n, 0ull);
        auto body = [&](auto kt) {
            using KT = decltype(kt);
            std::vector<KT> keys(kbase[n] + 16, (KT)0x5A5A5A5A5A5A5A5Aull);
            auto tile_pass = [&](auto modec) {
                constexpr int M = decltype(modec)::value;
                emu::launch(grid, thr, 0, [&]() {
                    if (k > 16) { if (thr == 32) sparse_tile_kernel<M, KT, true, 32>(arena.data(), tiles.data(), cta_begin.data(), k, 0u, tile_hist.data(), keys.data(), kbase.data(), bshift);
                                  else sparse_tile_kernel<M, KT, true, 64>(arena.data(), tiles.data(), cta_begin.data(), k, 0u, tile_hist.data(), keys.data(), kbase.data(), bshift); }
                    else { if (thr == 32) sparse_tile_kernel<M, KT, false, 32>(arena.data(), tiles.data(), cta_begin.data(), k, 0u, tile_hist.data(), keys.data(), kbase.data(), bshift);
                           else sparse_tile_kernel<M, KT, false, 64>(arena.data(), tiles.data(), cta_begin.data(), k, 0u, tile_hist.data(), keys.data(), kbase.data(), bshift); }
                });
            };
            tile_pass(std::integral_constant<int, 0>());
            emu::launch(n, 1024, 0, [&]() { sparse_tile_scan_kernel(tile_hist.data(), file_t0.data(), boff.data(), totals.data(), 0u, path16 ? 7u : 0u, path16 ? tile_place.data() : (uint32_t *)nullptr); });
            std::vector<uint16_t> keys16(path16 ? kbase[n] + 64 : 0, (uint16_t)0x5A5A);
            if (path16) emu::launch(grid, thr, sp16_scatter_smem(), [&]() {
                if (thr == 32) sparse_wc_scatter_kernel<32>(arena.data(), tiles.data(), cta_begin.data(), k, 0u, tile_hist.data(), tile_place.data(), keys16.data(), kbase.data());
                else sparse_wc_scatter_kernel<64>(arena.data(), tiles.data(), cta_begin.data(), k, 0u, tile_hist.data(), tile_place.data(), keys16.data(), kbase.data()); });
            else tile_pass(std::integral_constant<int, 1>());
            std::vector<unsigned long long> codes;
            std::vector<uint32_t> cnts;
            std::vector<unsigned long long> row_off(n + 1, 0ull);
            if constexpr (sizeof(KT) == 4) {
                if (path16) {
                    emu::launch(n * nbk16, 64, 0, [&]() { sparse16_distinct_kernel(keys16.data(), kbase.data(), file_t0.data(), tile_hist.data(), tile_place.data(), nbk16, nd.data()); });
                    emu::launch(n, 1024, 0, [&]() { sparse_scan_distinct_kernel(nd.data(), ooff32.data(), nd_file.data()); });
                    unsigned long long run = 0;
                    for (int f = 0; f < n; f++) { out_base[f] = run; run += nd_file[f]; row_off[f + 1] = run; }
                    codes.assign(run + 1, 0ull); cnts.assign(run + 1, 0u);
                    emu::launch(2 * n * nbk16, 64, 16384 * 4, [&]() {
                        sparse16_emit_kernel<64>(keys16.data(), kbase.data(), file_t0.data(), tile_hist.data(), tile_place.data(), nbk16, ooff32.data(), out_base.data(), codes.data(), cnts.data()); });
                } else if (hist_path) {
                    emu::launch(n_items / SP_HIST_WARPS, 32 * SP_HIST_WARPS, 0, [&]() { sparse_bucket_distinct_kernel((const uint32_t *)keys.data(), kbase.data(), boff.data(), R, nd.data()); });
                    emu::launch(n, 1024, 0, [&]() { sparse_scan_distinct_kernel(nd.data(), ooff32.data(), nd_file.data()); });
                    unsigned long long run = 0;
                    for (int f = 0; f < n; f++) { out_base[f] = run; run += nd_file[f]; row_off[f + 1] = run; }
                    codes.assign(run + 1, 0ull); cnts.assign(run + 1, 0u);
                    emu::launch(n_items / SP_HIST_WARPS, 32 * SP_HIST_WARPS, (size_t)SP_HIST_WARPS * ((size_t)1 << R) * 4, [&]() {
                        sparse_bucket_emit_kernel((const uint32_t *)keys.data(), kbase.data(), boff.data(), R, ooff32.data(), out_base.data(), codes.data(), cnts.data()); });
                }
            }
            if (!hist_path) {
                emu::launch(n_items, 64, 65536, [&]() { sparse_sort_kernel<KT, 64>(keys.data(), kbase.data(), boff.data(), S, nd.data()); });
                emu::launch(1, 1024, 0, [&]() { sparse_scan_items_kernel(nd.data(), ooff.data(), n_items); });
                codes.assign(ooff[n_items] + 1, 0ull); cnts.assign(ooff[n_items] + 1, 0u);
                emu::launch(n_items, 64, 0, [&]() { sparse_emit_kernel<KT, 64>(keys.data(), kbase.data(), boff.data(), S, ooff.data(), codes.data(), cnts.data()); });
                for (int f = 0; f < n; f++) row_off[f + 1] = ooff[(size_t)(f + 1) * S];
            }
            for (int f = 0; f < n; f++) {
                printf("%llu", totals[f]);
                for (unsigned long long e = row_off[f]; e < row_off[f + 1]; e++) printf(" %llu:%u", codes[e], cnts[e]);
                printf("\nF\n");
            }
        };
        if (k <= 16) body((uint32_t)0); else body((unsigned long long)0);
        return 0;
    }
    size_t NB = (size_t)1 << (2 * k);
    if (mode == 2) {
        // partitioned shared-memory kernel + u32 fold, as kf_api.cu launches them for k = 8..10
        std::vector<int> file_t0(n + 1, 0);
        { size_t t = 0; for (int f = 0; f < n; f++) { file_t0[f] = (int)t; while (t < tiles.size() && tiles[t].file == (uint32_t)f) t++; } file_t0[n] = (int)t; }
        const int TBITS = k == 3 ? 0 : k == 4 ? 3 : 5;   // k = 3: one partition (the k = 8 shape), k = 4: 3 partitions of 96 bins, k = 5: 11
        // pass A: (file, 0) items parse the text once and write the decoded stream; pass B: (file, p >= 1) items count from it
        std::vector<uint32_t> items_a, items_b;
        for (int f = 0; f < n; f++) if (len[f] && arena[off[f]] == '>') {
            items_a.push_back((uint32_t)f << 8);
            for (uint32_t pp = 1; pp < (TBITS ? ((1u << TBITS) + 2) / 3 : 1u); pp++) items_b.push_back(((uint32_t)f << 8) | pp);
        }
        std::vector<uint32_t> fwd32((size_t)n * NB, 0u);
        std::vector<uint2> stream(arena.size() / CHUNK * 32 + 64, uint2{0xDEADBEEFu, 0u});   // garbage that would be counted if read unwritten
        unsigned int counter[2] = {0, 0};
        const int thr = threads == 512 ? 64 : threads;
#define RUN_PART(KK, PBB, TT, MODE, ITEMS, CNT) count_fasta_part_kernel<KK, PBB, TT, MODE>(arena.data(), tiles.data(), file_t0.data(), ITEMS.data(), (int)ITEMS.size(), fwd32.data(), 0u, CNT, stream.data())
        if (k == 4) {
            emu::launch(grid, thr, PartSink<4, 3>::NWORDS * 4, [&]() { if (thr == 32) RUN_PART(4, 3, 32, 1, items_a, &counter[0]); else RUN_PART(4, 3, 64, 1, items_a, &counter[0]); });
            emu::launch(grid, thr, PartSink<4, 3>::NWORDS * 4, [&]() { if (thr == 32) RUN_PART(4, 3, 32, 2, items_b, &counter[1]); else RUN_PART(4, 3, 64, 2, items_b, &counter[1]); });
        } else if (k == 5) {
            emu::launch(grid, thr, PartSink<5, 5>::NWORDS * 4, [&]() { if (thr == 32) RUN_PART(5, 5, 32, 1, items_a, &counter[0]); else RUN_PART(5, 5, 64, 1, items_a, &counter[0]); });
            emu::launch(grid, thr, PartSink<5, 5>::NWORDS * 4, [&]() { if (thr == 32) RUN_PART(5, 5, 32, 2, items_b, &counter[1]); else RUN_PART(5, 5, 64, 2, items_b, &counter[1]); });
        } else if (k == 3) {   // one partition: the k = 8 shape (MODE 0)
            emu::launch(grid, thr, PartSink<3, 0>::NWORDS * 4, [&]() { if (thr == 32) RUN_PART(3, 0, 32, 0, items_a, &counter[0]); else RUN_PART(3, 0, 64, 0, items_a, &counter[0]); });
        }
        else { fprintf(stderr, "partitioned emu: k = 3, 4 or 5\n"); return 2; }
        std::vector<uint32_t> canon; canonical_codes(k, canon);
        long long V = (long long)canon.size();
        std::vector<unsigned long long> counts((size_t)n * V), totals(n);
        std::vector<double> freq((size_t)n * V);
        emu::launch(n, 64, 0, [&]() { fold_normalize_kernel<uint32_t>(fwd32.data(), canon.data(), k, V, 0u, 0u, (const uint32_t *)nullptr, (const uint32_t *)nullptr, counts.data(), freq.data(), (float *)nullptr, totals.data()); });
        for (int f = 0; f < n; f++) {
            printf("%llu", totals[f]);
            for (long long i = 0; i < V; i++) printf(" %llu", counts[(size_t)f * V + i]);
            printf("\nF");
            for (long long i = 0; i < V; i++) printf(" %.17g", freq[(size_t)f * V + i]);
            printf("\n");
        }
        return 0;
    }
    // rows: same rule as kf_api.cu:build_rows with stride 1 (every emulated CTA is a line-kernel CTA)
    std::vector<uint32_t> file_row(n + 1, 0), file_first_cta(grid, 0);   // (second one: rank of each CTA for its first file)
    {
        std::vector<uint32_t> cnt(n, 0); std::vector<int> last(n, -1);
        if (k == 7)
            for (int b = 0; b < grid; b++) {
                if (cta_begin[b] < cta_begin[b + 1]) file_first_cta[b] = cnt[tiles[cta_begin[b]].file];
                for (int t = cta_begin[b]; t < cta_begin[b + 1]; t++) {
                    uint32_t f = tiles[t].file;
                    if (last[f] != b) { cnt[f]++; last[f] = b; }
                }
            }
        for (int f = 0; f < n; f++) file_row[f + 1] = file_row[f] + std::max(1u, cnt[f]);
    }
    std::vector<unsigned long long> fwd((size_t)file_row[n] * NB, 0xDEADBEEFCAFEull);   // garbage: every row must be written or zeroed by the kernels
    // same launch sequence as kf_api.cu:run_files -- probe, one line-grid launch per width, generic kernel
    std::vector<uint8_t> formats(n);
    for (int i = 0; i < n; i++) formats[i] = len[i] ? arena[off[i]] : 0;
    std::vector<uint32_t> file_P(n, 0);
    std::vector<uint32_t> wc(8, 0);
    const bool lg = use_lg && k == 7 && !fw;
    emu::launch(n, 32, 0, [&]() { probe_line_width_kernel(arena.data(), off.data(), len.data(), formats.data(), n, lg ? 0u : 1u, file_P.data(), wc.data(), fwd.data(), file_row.data(), (uint32_t)NB); });
    if (lg) {
        // one launch for all widths and the long-line files, as kf_api.cu does
        if (threads == 512) run_lg<0, 512>(arena.data(), tiles, cta_begin, grid, file_P.data(), off.data(), len.data(), fwd.data(), file_row.data(), file_first_cta.data(), wc.data());
        else if (threads == 64) run_lg<0, 64>(arena.data(), tiles, cta_begin, grid, file_P.data(), off.data(), len.data(), fwd.data(), file_row.data(), file_first_cta.data(), wc.data());
        else run_lg<0, 32>(arena.data(), tiles, cta_begin, grid, file_P.data(), off.data(), len.data(), fwd.data(), file_row.data(), file_first_cta.data(), wc.data());
        int nlg = 0, nvl = 0; for (auto P : file_P) { nlg += P != 0; nvl += P == KF_P_VIRTUAL; }
        fprintf(stderr, "linegrid files: %d of %d (virtual lines: %d)\n", nlg, n, nvl);
    }
#define RUN(KK) case KK: if (threads == 512) run<KK, 512>(arena.data(), tiles, cta_begin, grid, fw, fwd.data(), file_row.data(), file_P.data(), wc.data()); else if (threads == 64) run<KK, 64>(arena.data(), tiles, cta_begin, grid, fw, fwd.data(), file_row.data(), file_P.data(), wc.data()); else run<KK, 32>(arena.data(), tiles, cta_begin, grid, fw, fwd.data(), file_row.data(), file_P.data(), wc.data()); break;
    switch (k) { RUN(3) RUN(4) RUN(5) RUN(7) default: fprintf(stderr, "k not built in emu\n"); return 2; }
    // FASTQ files: same plan as kf_api.cu (tiles of 32 lane ranges; tile_chunks <= 64 shrinks them so that small test
    // files still span several tiles -- the lane ranges stay FQ_LANE_BYTES, the tile just ends early)
    {
        std::vector<Tile> fq;
        uint32_t fq_tile = tile_chunks >= 64 ? 32 * FQ_LANE_BYTES / CHUNK : tile_chunks * 8;
        for (int f = 0; f < n; f++) {
            if (!len[f] || arena[off[f]] != '@') continue;
            uint64_t fc0 = off[f] / CHUNK, nch = (len[f] + CHUNK - 1) / CHUNK;
            for (uint64_t pos = 0; pos < nch; pos += fq_tile)
                fq.push_back(Tile{(uint32_t)(fc0 + pos), (uint32_t)std::min<uint64_t>(fq_tile, nch - pos), (uint32_t)f, (uint32_t)fc0});
        }
        if (!fq.empty()) {
            std::vector<int> cb(grid + 1);
            for (int b = 0; b <= grid; b++) cb[b] = (int)((uint64_t)fq.size() * b / grid);
            std::vector<unsigned long long> err(n, ~0ull);
            size_t smem = sizeof(uint32_t) << (2 * k);
#define RUNQ(KK) case KK: if (threads == 64) emu::launch(grid, 64, smem, [&]() { count_fastq_smem_kernel<KK, 64, 1>(arena.data(), fq.data(), cb.data(), off.data(), len.data(), fwd.data(), file_row.data(), err.data()); }); \
                          else emu::launch(grid, 32, smem, [&]() { count_fastq_smem_kernel<KK, 32, 1>(arena.data(), fq.data(), cb.data(), off.data(), len.data(), fwd.data(), file_row.data(), err.data()); }); break;
            const size_t psmem = (32768 + 16384 + 2 * 2 + 4) * sizeof(uint32_t);
            if (k == 7) {   // 8-mer pair histogram, as kf_api.cu launches it for k = 7 (stride 1 here: every emulated CTA is one)
                if (threads == 64) emu::launch(grid, 64, psmem, [&]() { count_fastq_pairs_kernel<64>(arena.data(), fq.data(), cb.data(), 1, off.data(), len.data(), fwd.data(), file_row.data(), err.data()); });
                else emu::launch(grid, 32, psmem, [&]() { count_fastq_pairs_kernel<32>(arena.data(), fq.data(), cb.data(), 1, off.data(), len.data(), fwd.data(), file_row.data(), err.data()); });
            } else
            switch (k) { RUNQ(3) RUNQ(4) RUNQ(5) default: return 2; }
            // files out of 4-line layout: the exact front-to-back walk, one warp per file (as kf_api.cu launches it)
            emu::launch((unsigned)((n * 32 + 127) / 128), 128, 0, [&]() { fastq_multiline_kernel<unsigned long long>(arena.data(), off.data(), len.data(), formats.data(), 0u, (uint32_t)n, k, fwd.data(), file_row.data(), err.data()); });
            for (int f = 0; f < n; f++)
                if (err[f] != ~0ull && err[f] - off[f] < len[f]) fprintf(stderr, "fastq layout violation file %d at %llu\n", f, err[f] - off[f]);
        }
    }
    std::vector<uint32_t> canon; canonical_codes(k, canon);
    long long V = (long long)canon.size();
    std::vector<unsigned long long> counts((size_t)n * V), totals(n);
    std::vector<double> freq((size_t)n * V);
    emu::launch(n, FOLD_THREADS, NB * sizeof(unsigned long long), [&]() { fold_normalize_smem_kernel<unsigned long long>(fwd.data(), canon.data(), k, V, 0u, lg ? file_P.data() : (const uint32_t *)nullptr, file_row.data(), counts.data(), freq.data(), (float *)nullptr, totals.data()); });
    for (int f = 0; f < n; f++) {
        printf("%llu", totals[f]);
        for (long long i = 0; i < V; i++) printf(" %llu", counts[(size_t)f * V + i]);
        printf("\n");
        printf("F");
        for (long long i = 0; i < V; i++) printf(" %.17g", freq[(size_t)f * V + i]);
        printf("\n");
    }
    return 0;
}
