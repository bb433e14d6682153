// Cluster variant of the fused implicit column stage: a LOADER CTA and a SOLVER CTA on two SMs of one thread-block cluster.
//
// Why: the two-sided product-form Thomas solve of a tile takes 3400-4000 cycles for a warp that has an SM sub-partition to
// itself, but 6800 cycles inside the one-CTA-per-tile kernel and 10^5 next to warps that stream global memory with LDG/STG --
// the solver's LDS/STS queue up behind those in the memory-instruction path -- while bulk-TMA traffic does not disturb it
// (scratch/solve_noise.cu).  Here no LDG/STG runs on the SM that solves:
//
//   loader CTA (cluster rank 0), 4 groups of 128 threads, group s serves slot s:
//        LDG.128 base + tendencies -> U0 = K1(...) in registers -> K3 residual through warp shuffles
//        -> wait until slot s of the solver CTA is free -> st.shared::cluster (distributed shared memory) of U0, the residual and
//        K/dz^2 into the solver CTA's slot -> remote mbarrier arrive (release.cluster) on full[s]
//   solver CTA (cluster rank 1):
//        solver warp s (CTA warps 0..3 = one per SM sub-partition): wait full[s] (+ W(s) by bulk TMA) -> solve in place ->
//        bar.arrive solved[s]
//        data group g (128 threads, slots g and g+2): bar.sync solved[s] -> U = U0 - dx, T_imp = (U - U0)/dtγ in place ->
//        bulk-TMA stores -> slot free: remote arrive on the loader's empty[s], bulk-TMA request of the slot's next W
//
// Tile i of a cluster lives in slot i % 4 and is handled by loader group / solver warp / data group of that slot, so every
// waiter meets the phases of its barriers in sequence.  Same device functions, operation order and roundings as the
// kernels of ops.cu: bit-identical results.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "cts_internal.h"
#include "fused_args.cuh"

using namespace cts_tile;
using cts_ew::axpy_combine;
using cts_ew::axpy_combine_n1;
using cts_ew::Vec;
using cts_ew::vec_pack;
using cts_ew::vec_unpack;
using cts_fused::FusedArgs;
using cts_fused::GroupSplit;
using cts_fused::timp_linear;

namespace {

constexpr int kSlots = 4;
constexpr int kGroup = 128;                 // threads of a loader group / a data group
constexpr int kThreads = kSlots * kGroup;   // 512 in both CTAs
constexpr int kHeader = 256;                // mbarriers
constexpr int kRounds = 4;                  // 16-byte vectors per thread and tile (512 per tile)
constexpr int kFlat = kRounds * kGroup * 16;  // 8 KiB: one flat tile

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_rank(const void* local_smem, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(local_smem)), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_v2(uint32_t raddr, double2 v) {
  asm volatile("st.shared::cluster.v2.f64 [%0], {%1, %2};" ::"r"(raddr), "d"(v.x), "d"(v.y) : "memory");
}
__device__ __forceinline__ void st_cluster_f64(uint32_t raddr, double v) {
  asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(raddr), "d"(v) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t raddr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
// wait with cluster-scope acquire and a watchdog (a protocol error must surface as a trap, never as a hung GPU)
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity, int tag) {
  uint32_t ok;
  long long t0 = 0;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (!ok) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ll) {
        printf("cts: fused-stage cluster watchdog: block %d thread %d barrier tag %d parity %u\n", (int)blockIdx.x, (int)threadIdx.x, tag, parity);
        __trap();
      }
    }
  } while (!ok);
}

template <int NT, bool MATRIX_FREE, int N1SPEC>
__global__ void __launch_bounds__(kThreads, 1) column_fused_stage_cluster_kernel(const FusedArgs<NT> a) {
  using T = double;
  using V = double2;
  constexpr int VN = 2;
  constexpr int NTS = NT > 0 ? NT : 1;
  constexpr int kCols = 16;
  extern __shared__ __align__(128) char smem[];
  const int pitch = a.pitch;
  const int tile_bytes = kCols * pitch;
  const int slot_bytes = (MATRIX_FREE ? 2 : 4) * tile_bytes + kFlat + (MATRIX_FREE ? 3 * kFlat : 0) + 128;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);       // [4] solver CTA: U0/residual of slot s have arrived (128 remote arrivals)
  uint64_t* wfull = full + kSlots;                          // [4] solver CTA: W of slot s has landed (TMA)
  uint64_t* empty = full + 2 * kSlots;                      // [4] loader CTA: slot s of the solver CTA is free (1 remote arrival)
  auto slot_base = [&](int s) { return smem + kHeader + (size_t)s * slot_bytes; };
  // slot layout: x (padded) | W tiles or cp scratch (padded) | U0 (flat) | [W out: dl|d|du flat] | K/dz^2 [16]

  const int nlev = a.nlev;
  const int col_bytes = nlev * (int)sizeof(T);
  const int cpc = col_bytes / 16;  // 32
  const int cshift = 31 - __clz(cpc);
  const uint32_t rank = cluster_ctarank();
  const size_t nclusters = gridDim.x / 2, cluster_id = blockIdx.x / 2;
  const size_t ntiles = (a.ncols + kCols - 1) / kCols;
  const int count = ntiles > cluster_id ? (int)((ntiles - cluster_id + nclusters - 1) / nclusters) : 0;  // tiles of this cluster
  auto tile_of = [&](int i, size_t& col0, int& nvalid) {
    const size_t tile = cluster_id + (size_t)i * nclusters;
    col0 = tile * kCols;
    nvalid = (int)((a.ncols - col0) < (size_t)kCols ? (a.ncols - col0) : (size_t)kCols);
  };
  if (threadIdx.x == 0) {
    for (int s = 0; s < kSlots; ++s) {
      mbar_init(full + s, kGroup);
      mbar_init(wfull + s, 1);
      mbar_init(empty + s, 1);
    }
    fence_mbar_init();
  }
  cluster_sync_all();  // both CTAs' barriers exist before anybody arrives remotely

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long acc[4] = {0, 0, 0, 0};
  const long long t_begin = a.dbg ? clock64() : 0;

  if (rank == 0) {
    // ================================================================== loader CTA ==============================
    const int s = threadIdx.x / kGroup;   // slot served by this group
    const int dt = threadIdx.x % kGroup;
    char* sl = slot_base(s);
    const uint32_t r_x = map_to_rank(sl, 1);
    const uint32_t r_U0 = map_to_rank(sl + (MATRIX_FREE ? 2 : 4) * tile_bytes, 1);
    const uint32_t r_K = map_to_rank(sl + (MATRIX_FREE ? 2 : 4) * tile_bytes + kFlat + (MATRIX_FREE ? 3 * kFlat : 0), 1);
    const uint32_t r_full = map_to_rank(full + s, 1);
    for (int j = 0, i = s; i < count; ++j, i += kSlots) {
      size_t col0;
      int nvalid;
      tile_of(i, col0, nvalid);
      const int total = nvalid * cpc;
      const size_t gvec0 = col0 * (size_t)nlev * sizeof(T) / 16;
      V U0r[kRounds], rr_v[kRounds];
      T acol[kRounds];
      constexpr int UN = 2;
#pragma unroll
      for (int r0 = 0; r0 < kRounds; r0 += UN) {
        V vb[UN], vt[NTS][UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) {
          const int q = dt + (r0 + u) * kGroup;
          vb[u] = V{};
#pragma unroll
          for (int t = 0; t < NT; ++t) vt[t][u] = V{};
          acol[r0 + u] = T(0);
          if (q < total) {
            const size_t gi = gvec0 + (size_t)q;
            vb[u] = reinterpret_cast<const V*>(a.ax.base)[gi];
#pragma unroll
            for (int t = 0; t < NT; ++t) vt[t][u] = reinterpret_cast<const V*>(a.ax.terms[t])[gi];
            acol[r0 + u] = ((const T*)a.Kcol)[col0 + (q >> cshift)];
          }
        }
#pragma unroll
        for (int u = 0; u < UN; ++u) {
          T eb[VN], uu[VN], rr[VN], et[NTS][VN];
          vec_unpack<T>(vb[u], eb);
#pragma unroll
          for (int t = 0; t < NT; ++t) vec_unpack<T>(vt[t][u], et[t]);
#pragma unroll
          for (int e = 0; e < VN; ++e) {
            T x[NTS];
#pragma unroll
            for (int t = 0; t < NT; ++t) x[t] = et[t][e];
            constexpr int N1 = GroupSplit<NT, N1SPEC>::value;
            if constexpr (N1 >= 0)
              uu[e] = axpy_combine_n1<T, T, NT, (N1 >= 0 ? N1 : 0)>(eb[e], x, a.ax);
            else
              uu[e] = axpy_combine<T, T, NT>(eb[e], x, a.ax);
          }
          U0r[r0 + u] = vec_pack<T>(uu);
          const int q = dt + (r0 + u) * kGroup;
          const int kk = q & (cpc - 1);
          T below = __shfl_up_sync(0xffffffffu, uu[VN - 1], 1, cpc);
          T above = __shfl_down_sync(0xffffffffu, uu[0], 1, cpc);
          if (kk == 0) below = T(0);
          if (kk == cpc - 1) above = T(0);
#pragma unroll
          for (int e = 0; e < VN; ++e) {
            const T um = e > 0 ? uu[e > 0 ? e - 1 : 0] : below;
            const T up = e + 1 < VN ? uu[e + 1 < VN ? e + 1 : e] : above;
            const T ti = timp_linear<T>(acol[r0 + u], um, uu[e], up);
            rr[e] = (T)__dsub_rn(__dadd_rn((double)uu[e], __dmul_rn(a.dtg, (double)ti)), (double)uu[e]);
          }
          rr_v[r0 + u] = vec_pack<T>(rr);
        }
      }
      // the solver CTA's slot is free (its previous tile has left by bulk stores)
      {
        const long long t0 = a.dbg ? clock64() : 0;
        mbar_wait_cluster(empty + s, (j & 1) ^ 1, 1);
        if (a.dbg && dt == 0) acc[1] += clock64() - t0;
      }
#pragma unroll
      for (int r = 0; r < kRounds; ++r) {
        const int q = dt + r * kGroup;
        if (q >= total) continue;
        const int col = q >> cshift, kk = q & (cpc - 1);
        st_cluster_v2(r_U0 + (uint32_t)q * 16u, U0r[r]);
        st_cluster_v2(r_x + (uint32_t)(col * pitch + kk * 16), rr_v[r]);
        if (a.temp) reinterpret_cast<V*>(a.temp)[gvec0 + q] = U0r[r];
        if (kk == 0) st_cluster_f64(r_K + (uint32_t)col * 8u, acol[r]);
      }
      mbar_arrive_remote(r_full);  // release.cluster: the stores above are visible to the solver CTA's acquire
    }
  } else {
    // ================================================================== solver CTA ==============================
    if (warp < kSlots) {
      // ---------------------------------------------------------------- solver warp = slot
      const int s = warp;
      char* sl = slot_base(s);
      char* sx = sl;
      char* sw = sl + tile_bytes;
      const T* sK = reinterpret_cast<const T*>(sl + (MATRIX_FREE ? 2 : 4) * tile_bytes + kFlat + (MATRIX_FREE ? 3 * kFlat : 0));
      const int col = lane_column(lane);
      const bool is_up = lane_is_up(lane);
      const int col_off = col * pitch;
      const int bar_solved = 1 + s;
      for (int j = 0, i = s; i < count; ++j, i += kSlots) {
        size_t col0;
        int nvalid;
        tile_of(i, col0, nvalid);
        const bool valid = col < nvalid;
        {
          const long long t0 = a.dbg ? clock64() : 0;
          mbar_wait_cluster(full + s, j & 1, 2);
          if (!MATRIX_FREE) mbar_wait_cluster(wfull + s, j & 1, 3);
          if (a.dbg && lane == 0) acc[1] += clock64() - t0;
        }
        const long long ts = a.dbg ? clock64() : 0;
        if (MATRIX_FREE) {
          const T ac = valid ? sK[col] : T(0);
          const double o = __dmul_rn(a.dtg, (double)ac);
          ConstRows<T> rows{(T)o, (T)__dsub_rn(__dmul_rn(-2.0, o), 1.0), sw};
          babe_solve_lanepair<T, ConstRows<T>, RhsFromTile<T>, true>(rows, RhsFromTile<T>{}, sx, col_off, nlev, is_up, valid);
        } else {
          StoredRows<T> rows{sw, sw + tile_bytes, sw + 2 * tile_bytes};
          babe_solve_lanepair<T, StoredRows<T>, RhsFromTile<T>, true>(rows, RhsFromTile<T>{}, sx, col_off, nlev, is_up, valid);
          fence_proxy_async_smem();  // the factors overwrote dl/du in place; the next writer of these bytes is a bulk copy
        }
        if (a.dbg && lane == 0) acc[2] += clock64() - ts;
        asm volatile("bar.arrive %0, %1;" ::"r"(bar_solved), "n"(32 + kGroup) : "memory");
      }
    } else if (warp < kSlots + 2 * (kGroup / 32)) {
      // ---------------------------------------------------------------- data group g: slots g and g + 2
      const int g = (warp - kSlots) / (kGroup / 32);
      const int dt = ((int)threadIdx.x - kSlots * 32) % kGroup;
      const int bar_group = 1 + kSlots + g;
      const DivConst dc = div_const_prepare(a.dtg);
      const bool emit = MATRIX_FREE && a.emit_w;
      auto issue_w = [&](int i) {  // first warp of the group: W(i) -> slot i % 4, one 1-D bulk copy per column and diagonal
        size_t col0;
        int nvalid;
        tile_of(i, col0, nvalid);
        const int s = i % kSlots;
        char* sw = slot_base(s) + tile_bytes;
        const size_t goff = col0 * (size_t)nlev;
        if (lane == 0) mbar_arrive_expect_tx(wfull + s, (uint32_t)(3 * nvalid * col_bytes));
        __syncwarp();
        for (int c = lane; c < nvalid; c += 32) {
          const size_t gg = goff + (size_t)c * nlev;
          bulk_g2s(sw + c * pitch, (const T*)a.dl + gg, col_bytes, wfull + s);
          bulk_g2s(sw + tile_bytes + c * pitch, (const T*)a.d + gg, col_bytes, wfull + s);
          bulk_g2s(sw + 2 * tile_bytes + c * pitch, (const T*)a.du + gg, col_bytes, wfull + s);
        }
      };
      if (!MATRIX_FREE && dt < 32) {
        if (g < count) issue_w(g);
        if (g + 2 < count) issue_w(g + 2);
      }
      for (int i = g; i < count; i += 2) {
        size_t col0;
        int nvalid;
        tile_of(i, col0, nvalid);
        const int s = i % kSlots;
        const int total = nvalid * cpc;
        const size_t goff = col0 * (size_t)nlev;
        char* sl = slot_base(s);
        char* sx = sl;
        char* sU0c = sl + (MATRIX_FREE ? 2 : 4) * tile_bytes;
        V* sU0 = reinterpret_cast<V*>(sU0c);
        char* sWo = sU0c + kFlat;
        const T* sK = reinterpret_cast<const T*>(sWo + (MATRIX_FREE ? 3 * kFlat : 0));
        {
          const long long t0 = a.dbg ? clock64() : 0;
          asm volatile("bar.sync %0, %1;" ::"r"(1 + s), "n"(32 + kGroup) : "memory");  // dx(i) is in sx
          if (a.dbg && dt == 0) acc[1] += clock64() - t0;
        }
        // epilogue in place: sU0 <- U = U0 - dx ; sx <- T_imp = (U - U0)/dtγ
#pragma unroll
        for (int r = 0; r < kRounds; ++r) {
          const int q = dt + r * kGroup;
          if (q >= total) continue;
          const int col = q >> cshift, kk = q & (cpc - 1);
          V* px = reinterpret_cast<V*>(sx + col * pitch + kk * 16);
          T e0[VN], ex[VN], eU[VN], eT[VN];
          vec_unpack<T>(sU0[q], e0);
          vec_unpack<T>(*px, ex);
#pragma unroll
          for (int e = 0; e < VN; ++e) {
            eU[e] = rsub(e0[e], ex[e]);
            eT[e] = (T)div_const((double)rsub(eU[e], e0[e]), dc);
          }
          sU0[q] = vec_pack<T>(eU);
          *px = vec_pack<T>(eT);
          if (emit) {  // W = dtγJ - I of the stored copy (same expressions as column_wfact_kernel)
            const double o = __dmul_rn(a.dtg, (double)sK[col]);
            T eo[VN], ed[VN];
#pragma unroll
            for (int e = 0; e < VN; ++e) {
              eo[e] = (T)o;
              ed[e] = (T)__dsub_rn(__dmul_rn(-2.0, o), 1.0);
            }
            const V vo = vec_pack<T>(eo);
            reinterpret_cast<V*>(sWo)[q] = vo;
            reinterpret_cast<V*>(sWo + kFlat)[q] = vec_pack<T>(ed);
            reinterpret_cast<V*>(sWo + 2 * kFlat)[q] = vo;
          }
        }
        fence_proxy_async_smem();  // generic-proxy writes -> visible to the bulk stores
        asm volatile("bar.sync %0, %1;" ::"r"(bar_group), "n"(kGroup) : "memory");
        {
          const long long t0 = a.dbg ? clock64() : 0;
          const uint32_t flat = (uint32_t)(nvalid * col_bytes);
          bool issued = false;
          if (dt == 0) {
            bulk_s2g((T*)a.U + goff, sU0c, flat);
            if (emit) {
              bulk_s2g((T*)const_cast<void*>(a.dl) + goff, sWo, flat);
              bulk_s2g((T*)const_cast<void*>(a.d) + goff, sWo + kFlat, flat);
              bulk_s2g((T*)const_cast<void*>(a.du) + goff, sWo + 2 * kFlat, flat);
            }
            issued = true;
          }
          if (a.Timp && dt < nvalid) {
            bulk_s2g((T*)a.Timp + goff + (size_t)dt * nlev, sx + dt * pitch, col_bytes);
            issued = true;
          }
          if (issued) {
            bulk_commit();
            bulk_wait_read0();  // the stores have read shared memory: the slot may be rewritten
          }
          asm volatile("bar.sync %0, %1;" ::"r"(bar_group), "n"(kGroup) : "memory");
          if (a.dbg && dt == 0) acc[2] += clock64() - t0;
        }
        // slot s is free: tell the loader group of this slot, and request the slot's next W
        if (dt == 0) mbar_arrive_remote(map_to_rank(empty + s, 0));
        if (!MATRIX_FREE && dt < 32 && i + kSlots < count) issue_w(i + kSlots);
      }
    }
  }
  if (a.dbg) {
    const bool rep = (rank == 0 && threadIdx.x == 0) || (rank == 1 && (threadIdx.x == 0 || threadIdx.x == kSlots * 32));
    if (rep) {
      acc[0] = clock64() - t_begin;
      const int role = rank == 0 ? 0 : (threadIdx.x == 0 ? 1 : 2);
      for (int k = 0; k < 4; ++k) a.dbg[((size_t)(blockIdx.x / 2) * 3 + role) * 4 + k] = acc[k];
    }
  }
  cluster_sync_all();  // nobody leaves while its peer may still touch its shared memory
}

template <int NT, bool MF, int N1SPEC>
int launch_cluster(cts_ctx* ctx, const FusedArgs<NT>& a_in, bool timeline) {
  FusedArgs<NT> a = a_in;
  const int tile_bytes = 16 * a.pitch;
  const size_t smem = kHeader + (size_t)kSlots * ((size_t)(MF ? 2 : 4) * tile_bytes + kFlat + (MF ? 3 * kFlat : 0) + 128);
  auto kern = column_fused_stage_cluster_kernel<NT, MF, N1SPEC>;
  CTS_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = ctx->stream;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cfg.gridDim = dim3(2 * (unsigned)(ctx->num_sms / 2));
  int max_clusters = 0;
  if (cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg) != cudaSuccess || max_clusters < 1) {
    cudaGetLastError();
    return CTS_EUNSUPPORTED;
  }
  const size_t ntiles = (a.ncols + 15) / 16;
  const size_t nclusters = std::min<size_t>({(size_t)max_clusters, (size_t)(ctx->num_sms / 2), ntiles});
  cfg.gridDim = dim3((unsigned)(2 * nclusters));
  a.dbg = nullptr;
  if (timeline) {
    CTS_CUDA(ctx, cudaMalloc(&a.dbg, nclusters * 12 * sizeof(long long)));
    CTS_CUDA(ctx, cudaMemsetAsync(a.dbg, 0, nclusters * 12 * sizeof(long long), ctx->stream));
  }
  CTS_CUDA(ctx, cudaLaunchKernelEx(&cfg, kern, a));
  ctx->launches++;
  if (timeline) {
    std::vector<long long> h(nclusters * 12);
    cudaStreamSynchronize(ctx->stream);
    cudaMemcpy(h.data(), a.dbg, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    cudaFree(a.dbg);
    static const char* roles[3] = {"loader group 0", "solver warp 0", "data group 0"};
    static const char* t1[3] = {"wait slot free", "wait full/W", "wait solved"};
    static const char* t2[3] = {"-", "SOLVE", "store+drain"};
    fprintf(stderr, "[cts cluster timeline] NT=%d mf=%d tiles=%zu clusters=%zu (max active %d) smem=%zu\n", NT, (int)MF, ntiles, nclusters, max_clusters, smem);
    for (int r = 0; r < 3; ++r) {
      double sum[4] = {0, 0, 0, 0};
      for (size_t c = 0; c < nclusters; ++c)
        for (int k = 0; k < 4; ++k) sum[k] += (double)h[(c * 3 + r) * 4 + k];
      fprintf(stderr, "  %-15s total %.0f cycles (%.0f per tile of the cluster); %s %.1f%% %s %.1f%%\n", roles[r], sum[0] / nclusters,
              sum[0] / nclusters / ((double)ntiles / nclusters), t1[r], 100 * sum[1] / sum[0], t2[r], 100 * sum[2] / sum[0]);
    }
  }
  return CTS_OK;
}

}  // namespace

int cts_fused_cluster_f64(cts_ctx* ctx, int nt, int n1spec, bool mf, const void* args, bool timeline) {
#define CASE(NT_, MF_, SPEC_) \
  if (nt == NT_ && mf == MF_ && n1spec == SPEC_) return launch_cluster<NT_, MF_, SPEC_>(ctx, *static_cast<const FusedArgs<NT_>*>(args), timeline);
  // the stages of a 4-stage IMEX-ARK step with dense lower triangles (ARS343): 1 / 3 / 5 operands
  CASE(1, true, 1) CASE(1, false, 1) CASE(3, true, 2) CASE(3, false, 2) CASE(5, true, 2) CASE(5, false, 2)
  CASE(2, true, 2) CASE(2, false, 2) CASE(4, true, 2) CASE(4, false, 2)
#undef CASE
  return CTS_EUNSUPPORTED;
}
