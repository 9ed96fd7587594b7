// jwc_dwt_fast.cu -- fused multi-level decimated filter bank on shared-memory tiles: FWT pyramid and WPT full tree.
//
// Reference (paths relative to /root/reference/src/main/java/jwave/):
//   one analysis step   transforms/wavelets/Wavelet.java:236-260   lo[i] = sum_j x[(2i+j) mod h] s[j], hi likewise
//   one synthesis step  transforms/wavelets/Wavelet.java:277-303   out[(2i+j) mod h] += c[i] sR[j] + c[i+h/2] wR[j]
//   FWT level loops     transforms/FastWaveletTransform.java:85-99, 133-151   (prefix h = N, N/2, ...)
//   WPT level loops     transforms/WaveletPacketTransform.java:98-120, 167-187 (every block of length h)
//
// A pass carries one tile of one node k levels down (forward) or up (inverse) without leaving shared memory; see
// jwc_dwt_plan.cuh for the geometry.  HBM traffic per pass: the tile is read once (1-D bulk TMA copies, periodic
// wrap = the copy is split at the node boundary) and every coefficient is written once (bulk stores).
//
// Inner loops: one work item = R consecutive outputs of one node.
//   analysis:  reads R + L/2 - 1 aligned sample pairs (LDS.128), 2*R*L DFMA -> R low-pass + R high-pass outputs
//   synthesis: reads R + L/2 - 1 (lo, hi) coefficient pairs (2 x LDS.64), 2*R*L DFMA -> R output pairs (STS.128)
// R is odd, so the lane stride (R x 16 B resp. R x 8 B) maps the lanes of every quarter/half warp to distinct banks.
#include <cstdlib>
#include <memory>

#include "jwc_dwt_plan.cuh"
#include "jwc_internal.cuh"
#include "jwc_tma.cuh"

namespace jwc {

namespace {

#ifndef JWC_SYN_DIRECT
#define JWC_SYN_DIRECT 0
#endif
constexpr bool kSynDirect = JWC_SYN_DIRECT != 0;   // experiment: constant-operand taps in the pyramid inverse for 12 <= L <= 20;
                                                   // measured (C3 db8 inverse): 7.98 ms vs 4.87 ms with the shared-memory taps -> off
constexpr int kUniformTapsMaxDwt = 10;   // longer filters read their taps from a shared-memory copy (see MODWT kernel)

void debug_dwt_plan(const char* what, const DwtPlan& plan, int64_t n, int levels, int L, bool tree) {
  static const bool on = getenv("JWC_DEBUG") != nullptr;
  if (!on) return;
  fprintf(stderr, "[jwc] %s %s n=%lld levels=%d L=%d:", tree ? "wpt" : "fwt", what, (long long)n, levels, L);
  for (const DwtPass& p : plan.passes)
    fprintf(stderr, " [l0=%d k=%d T=%d cap=%d mode=%d smem=%zu thr=%d]", p.l0, p.k, p.T, p.cap, p.mode, p.smem, p.threads);
  fprintf(stderr, "\n");
}

struct DwtPassArgs {
  const double* in;    // forward: nodes at depth l0; inverse: coefficient source of the children (see below)
  const double* ain;   // inverse FWT only: A_{l0+k}
  double* out;         // forward: D / leaf destination (final layout); inverse: depth-l0 node destination
  double* aout;        // forward FWT only: A_{l0+k} destination
  int64_t in_sig, ain_sig, out_sig, aout_sig;   // signal strides
  int64_t N;           // full signal length
  int64_t h;           // node length at depth l0 (= N >> l0)
  int l0, k, T, tiles, nodes, cap, mode;
  int pf_dist;        // L2 prefetch distance in CTAs (0 = off)
  int top_barrier;    // inverse: 1 = round-1 form of the tile wait (see the level loop)
  int upfront;        // pyramid inverse, bulk mode: every D tile of the pass is requested in the prologue (own slot, own mbarrier)
  int log_tiles;      // tiles and nodes are powers of two: the CTA index is taken apart with shifts (a 64-bit
                      // division is a ~100-instruction subroutine, and a CTA only lives for a few thousand)
  unsigned nblocks;
  int hl[16];             // inverse: left halo of the depth-jj arrays (dwt_inv_halo), precomputed on the host
  // split-device mode (one long series in chunks on several devices, batch = 1): the arrays this pass reads are chunks
  // of longer ones, so what lies past their end (forward) / before their start (inverse) comes from `halo`, a small
  // buffer filled with the ring neighbour's boundary samples (cudaMemcpyPeerAsync), not from the periodic wrap.
  //   forward: halo[node * halo_stride + i] = sample h + i of the node            (right neighbour's head)
  //   inverse: slot s (halo_stride doubles) holds the samples before position 0, right-aligned (left neighbour's
  //            tail); tree: slot = node * 2^k + leaf; pyramid: slot 0 = A_{l0+k}, slot 1 + (k - jj) = D_{l0+jj}
  const double* halo;     // nullptr = ordinary periodic transform
  int64_t halo_stride;
};

// bulk load of `len` doubles starting at node position `start` of a node of length hn at `base` into dst.  One thread;
// the mbarrier must already expect the bytes.  start, hn, len are even.
//   hi == nullptr && lo_end == nullptr: circular (any sign of start, any number of wraps)
//   split-device mode: positions >= hn continue at hi[pos - hn], positions < 0 come from lo_end[pos] (single overhang)
__device__ __forceinline__ void bulk_load_circ(double* dst, const double* base, int64_t start, int len, int64_t hn,
                                               uint64_t* bar, const double* hi = nullptr, const double* lo_end = nullptr) {
  const bool circ = (hi == nullptr && lo_end == nullptr);
  // interior tile (all but the first / last of a node): one copy, no wrap logic.  The general path below costs the issuing
  // thread 100-130 instructions per tile (ptxas turns its subtract loops into a division) -- with k + 1 tiles per CTA
  // that was a serial prefix of 300-600 instructions in front of the CTA's first byte (profiles/r2_ncu_fwt_inverse.txt)
  if (circ && start >= 0 && start + len <= hn) {
    ptx::bulk_g2s(dst, base + start, (uint32_t)len * 8u, bar);
    return;
  }
  // no 64-bit modulo here: it would be a subroutine call inside the level loop of the inverse kernel, and everything
  // live across that call (accumulator / tap registers of the other threads' code path) pays for it.  |start| is at
  // most a halo (tens of samples), so the two loops run a handful of times for one thread.
  int64_t pos = start;
  if (circ) {
    while (pos < 0) pos += hn;
    while (pos >= hn) pos -= hn;
  }
  int done = 0;
  while (done < len) {
    const double* src;
    int64_t run;
    if (pos < 0) { run = -pos; src = lo_end + pos; }
    else if (pos >= hn) {
      if (circ) { pos = 0; continue; }
      run = len - done; src = hi + (pos - hn);
    } else { run = hn - pos; src = base + pos; }
    if (run > len - done) run = len - done;
    ptx::bulk_g2s(dst + done, src, (uint32_t)run * 8u, bar);
    done += (int)run;
    pos += run;
  }
}

__device__ __forceinline__ void scalar_load_circ(double* dst, const double* base, int64_t start, int len, int64_t hn,
                                                 int tid, int nt, const double* hi = nullptr,
                                                 const double* lo_end = nullptr) {
  if (hi != nullptr || lo_end != nullptr) {   // split-device mode
    for (int e = tid; e < len; e += nt) {
      const int64_t pos = start + e;
      dst[e] = pos < 0 ? lo_end[pos] : (pos >= hn ? hi[pos - hn] : base[pos]);
    }
    return;
  }
  // node lengths are < 2^31 (checked by the callers): 32-bit remainder, no division subroutine
  const int h32 = (int)hn;
  int s32 = (int)(start % hn);
  if (s32 < 0) s32 += h32;
  for (int e = tid; e < len; e += nt) {
    const int pos = (int)(((unsigned)s32 + (unsigned)e) % (unsigned)h32);
    dst[e] = base[pos];
  }
}

// Taps for long filters come from the kernel-parameter constant bank through an index the compiler cannot prove
// uniform: that makes every tap one LDC.64 into an ordinary register.  (Uniform-register operands run out at 2L > 31
// taps and ptxas then shuttles every tap through LDC + R2UR; a shared-memory copy read with broadcast LDS.128 costs a
// full 4-wavefront LSU slot per load -- measured: 48 % of all shared-memory wavefronts of the db8 FWT kernel.)
__device__ __forceinline__ const double* const_taps(const FilterPair& f) {
  int z;
  asm volatile("mov.u32 %0, 0;" : "=r"(z));
  return reinterpret_cast<const double*>(&f) + z;   // f0 at [0, 64), f1 at [64, 128)
}

// =========================================================================================================================
// forward (analysis)
// =========================================================================================================================
// QMF variant: when the high-pass filter is the exact quadrature mirror of the low-pass one (every orthogonal wavelet
// of the reference, wavelets/Wavelet.java:104-122: w[j] = (-1)^j s[L-1-j]) and 12 <= L <= 20, the L low-pass taps stay
// RESIDENT in registers for the whole kernel (2L registers, fewer than the rolling window) and the high-pass taps are
// the same registers read mirrored with a free sign flip: no tap loads in the inner loop at all.
template <int L>
constexpr bool qmf_supported() { return L >= 12 && L <= 20; }

template <int L, int R, bool WITH_HI, bool QMF, int NS>
__device__ __forceinline__ void ana_item(const double2* __restrict__ px, const FilterPair& f, const double* __restrict__ taps,
                                         const double (&sreg)[NS], double (&lo)[R], double (&hi)[R]) {
  constexpr bool ST = (L > kUniformTapsMaxDwt) && !QMF;
  constexpr int HL = L / 2;
#pragma unroll
  for (int r = 0; r < R; r++) { lo[r] = 0.0; hi[r] = 0.0; }
  double ts[L], tw[L];
#pragma unroll
  for (int pp = 0; pp < R + HL - 1; ++pp) {
    if (ST && pp < HL) {   // taps 2pp, 2pp+1 are first needed by output 0 at pair pp
      ts[2 * pp] = taps[2 * pp];
      ts[2 * pp + 1] = taps[2 * pp + 1];
      if (WITH_HI) {
        tw[2 * pp] = taps[JWC_MAX_TAPS + 2 * pp];
        tw[2 * pp + 1] = taps[JWC_MAX_TAPS + 2 * pp + 1];
      }
    }
    const double2 x = px[pp];
#pragma unroll
    for (int r = 0; r < R; r++) {
      const int q = pp - r;   // tap pair index: j = 2q, 2q+1 (ascending j per output, the reference's order)
      if (q >= 0 && q < HL) {
        if (QMF) {
          lo[r] = fma(x.x, sreg[(2 * q) % NS], lo[r]);
          lo[r] = fma(x.y, sreg[(2 * q + 1) % NS], lo[r]);
          if (WITH_HI) {   // w[2q] = +s[L-1-2q], w[2q+1] = -s[L-2-2q]
            hi[r] = fma(x.x, sreg[(L - 1 - 2 * q) % NS], hi[r]);
            hi[r] = fma(x.y, -sreg[(L - 2 - 2 * q) % NS], hi[r]);
          }
        } else {
          lo[r] = fma(x.x, ST ? ts[2 * q] : f.f0[2 * q], lo[r]);
          lo[r] = fma(x.y, ST ? ts[2 * q + 1] : f.f0[2 * q + 1], lo[r]);
          if (WITH_HI) {
            hi[r] = fma(x.x, ST ? tw[2 * q] : f.f1[2 * q], hi[r]);
            hi[r] = fma(x.y, ST ? tw[2 * q + 1] : f.f1[2 * q + 1], hi[r]);
          }
        }
      }
    }
  }
}

template <int L, int R, bool TREE, bool QMF, int NS>
__device__ __forceinline__ void ana_level(double* smem, const FilterPair& f, const double (&sreg)[NS], int oT, int oin,
                                          int oout, int st_in, int st_out, int len_out, int own, int parents, int tid,
                                          int nt) {
  const int nb = (len_out + R - 1) / R;
  const int items = parents * nb;
  // tap source: constant bank for 10 < L <= 20, shared-memory copy above (ptxas spills the hoisted LDC results there)
  const double* ctaps = (L <= 20) ? const_taps(f) : (smem + oT);
#pragma unroll 1
  for (int w = tid; w < items; w += nt) {
    const int q = TREE ? (w / nb) : 0;
    const int i0 = (w - q * nb) * R;
    const double2* px = reinterpret_cast<const double2*>(smem + oin + q * st_in) + i0;
    double lo[R], hi[R];
    const int olo = oout + (2 * q) * st_out + i0, ohi = olo + st_out;
    const bool full = i0 + R <= len_out;
    if (!TREE && i0 >= own) {   // FWT: the halo part of D is produced by the neighbouring tile
      ana_item<L, R, false, QMF, NS>(px, f, ctaps, sreg, lo, hi);
      if (full) {
#pragma unroll
        for (int r = 0; r < R; r++) smem[olo + r] = lo[r];
      } else {
#pragma unroll
        for (int r = 0; r < R; r++)
          if (i0 + r < len_out) smem[olo + r] = lo[r];
      }
    } else {
      ana_item<L, R, true, QMF, NS>(px, f, ctaps, sreg, lo, hi);
      if (full) {
#pragma unroll
        for (int r = 0; r < R; r++) {
          smem[olo + r] = lo[r];
          smem[ohi + r] = hi[r];
        }
      } else {
#pragma unroll
        for (int r = 0; r < R; r++)
          if (i0 + r < len_out) {
            smem[olo + r] = lo[r];
            smem[ohi + r] = hi[r];
          }
      }
    }
  }
}

template <int L, int RMAX, bool TREE, bool QMF>
__global__ void __launch_bounds__(256, (L > 10 ? 2 : 3)) dwt_fwd_pass_kernel(const __grid_constant__ DwtPassArgs a,
                                                              const __grid_constant__ FilterPair f) {
  extern __shared__ __align__(128) double smem[];
  const int tid = threadIdx.x, nt = blockDim.x;
  constexpr int NS = QMF ? L : 1;
  double sreg[NS];
  if (QMF) {
    const double* ct = const_taps(f);
#pragma unroll
    for (int m = 0; m < NS; m++) sreg[m] = ct[m];
  } else {
    sreg[0] = 0.0;
  }
  const int oT = 2 * a.cap;                         // tap copy, then the mbarrier
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + oT + 2 * JWC_MAX_TAPS);
  if (L > 20) {   // shorter filters take their taps from the constant bank (uniform registers)
    for (int t = tid; t < JWC_MAX_TAPS; t += nt) {
      smem[oT + t] = f.f0[t];
      smem[oT + JWC_MAX_TAPS + t] = f.f1[t];
    }
  }
  const unsigned bid = blockIdx.x;
  const int ti = (int)(bid & (unsigned)(a.tiles - 1));
  const int p = (int)((bid >> a.log_tiles) & (unsigned)(a.nodes - 1));
  const int64_t b = (int64_t)(bid >> (a.log_tiles + a.l0 * (a.nodes > 1 ? 1 : 0)));
  const int tlen = (int)((a.h < a.T) ? a.h : a.T);
  const int64_t a0 = (int64_t)ti * tlen;
  const double* node = a.in + b * a.in_sig + (int64_t)p * a.h;
  const bool bulk = (a.mode == DWT_BULK);
  const int H = (L - 2) * ((1 << a.k) - 1);
  const double* fhalo = a.halo ? a.halo + (int64_t)p * a.halo_stride : nullptr;   // split-device mode: right neighbour's head

  // ---- load: node samples a0 .. a0 + tlen + H - 1 (periodic) ---------------------------------------------------------
  if (bulk) {
    if (tid == 0) {
      ptx::mbar_init(bar, 1);
      ptx::fence_mbar_init();
      ptx::mbar_expect_tx(bar, (uint32_t)(tlen + H) * 8u);
      bulk_load_circ(smem, node, a0, tlen + H, a.h, bar, fhalo);
      if (a.pf_dist > 0 && blockIdx.x + (unsigned)a.pf_dist < a.nblocks) {   // one wave ahead into L2
        const unsigned nb = blockIdx.x + (unsigned)a.pf_dist;
        const int ti2 = (int)(nb & (unsigned)(a.tiles - 1));
        const int p2 = (int)((nb >> a.log_tiles) & (unsigned)(a.nodes - 1));
        const int64_t b2 = (int64_t)(nb >> (a.log_tiles + a.l0 * (a.nodes > 1 ? 1 : 0)));
        ptx::bulk_prefetch_l2(a.in + b2 * a.in_sig + (int64_t)p2 * a.h + (int64_t)ti2 * tlen, (uint32_t)tlen * 8u);
      }
      ptx::mbar_wait(bar, 0);
    }
    __syncthreads();
    ptx::mbar_wait(bar, 0);
  } else {
    scalar_load_circ(smem, node, a0, tlen + H, a.h, tid, nt, fhalo);
    __syncthreads();
  }

  // ---- k levels ---------------------------------------------------------------------------------------------------------
  for (int jj = 1; jj <= a.k; jj++) {
    const int oin = (jj & 1) ? 0 : a.cap, oout = (jj & 1) ? a.cap : 0;
    const int halo_in = (L - 2) * ((1 << (a.k - jj + 1)) - 1), halo_out = (L - 2) * ((1 << (a.k - jj)) - 1);
    const int len_in = (tlen >> (jj - 1)) + halo_in, len_out = (tlen >> jj) + halo_out;
    const int st_in = len_in + (len_in & 1) + 2 * kDwtR, st_out = len_out + (len_out & 1) + 2 * kDwtR;
    const int own = tlen >> jj;                       // outputs of each child that belong to this tile
    const int parents = TREE ? (1 << (jj - 1)) : 1;
    // rows per item: as many as the build offers unless that leaves more than 3/4 of the threads without an item
    const int Rsel = dwt_pick_r(L, len_out, parents, nt);
    switch (Rsel) {
      case 7:
        if constexpr (RMAX >= 7) {
          ana_level<L, 7, TREE, QMF, NS>(smem, f, sreg, oT, oin, oout, st_in, st_out, len_out, own, parents, tid, nt);
          break;
        }
      case 5:
        if constexpr (RMAX == 5) {
          ana_level<L, 5, TREE, QMF, NS>(smem, f, sreg, oT, oin, oout, st_in, st_out, len_out, own, parents, tid, nt);
          break;
        }
      case 3:
        ana_level<L, 3, TREE, QMF, NS>(smem, f, sreg, oT, oin, oout, st_in, st_out, len_out, own, parents, tid, nt);
        break;
      default:
        ana_level<L, 1, TREE, QMF, NS>(smem, f, sreg, oT, oin, oout, st_in, st_out, len_out, own, parents, tid, nt);
    }

    // ---- ship what is final after this level ------------------------------------------------------------------------------
    const bool last = (jj == a.k);
    const bool vec_ok = bulk && (own & 1) == 0;
    if (vec_ok) ptx::fence_proxy_async();
    if (bulk && tid < 32) ptx::bulk_wait_read<0>();   // stores issued one level ago have finished reading this buffer pair
    __syncthreads();
    if (!TREE) {
      // D_{l0+jj}: owned prefix of the high-pass child -> out[(N >> (l0+jj)) + (a0 >> jj) ...]
      const double* src = smem + oout + st_out;
      double* dst = a.out + b * a.out_sig + (a.N >> (a.l0 + jj)) + (a0 >> jj);
      if (vec_ok) {
        if (tid == 0) {
          ptx::bulk_s2g(dst, src, (uint32_t)own * 8u);
          if (last) ptx::bulk_s2g(a.aout + b * a.aout_sig + (a0 >> jj), smem + oout, (uint32_t)own * 8u);
          ptx::bulk_commit();
        }
      } else {
        for (int e = tid; e < own; e += nt) dst[e] = src[e];
        if (last) {
          double* ad = a.aout + b * a.aout_sig + (a0 >> jj);
          for (int e = tid; e < own; e += nt) ad[e] = smem[oout + e];
        }
      }
    } else if (last) {
      // 2^k leaves in natural (Paley) order: leaf c of this node lives at node_base + c * (h >> k)
      const int leaves = 1 << a.k;
      double* dst0 = a.out + b * a.out_sig + (int64_t)p * a.h + (a0 >> a.k);
      const int64_t leaf_stride = a.h >> a.k;
      if (vec_ok) {
        if (tid < 32) {
          for (int c = tid; c < leaves; c += 32) ptx::bulk_s2g(dst0 + c * leaf_stride, smem + oout + c * st_out, (uint32_t)own * 8u);
          ptx::bulk_commit();
        }
      } else {
        for (int e = tid; e < leaves * own; e += nt) {
          const int c = e / own, i = e - c * own;
          dst0[c * leaf_stride + i] = smem[oout + c * st_out + i];
        }
      }
    }
  }
  if (bulk && tid < 32) ptx::bulk_wait_read<0>();
}


// =========================================================================================================================
// inverse (synthesis)
// =========================================================================================================================
// DIRECT: taps as constant-bank operands of the DFMAs (f.f0[static index]) also for long filters -- no tap registers
template <int L, int R, bool QMF, int NS, bool DIRECT = false>
__device__ __forceinline__ void syn_item(const double* __restrict__ plo, const double* __restrict__ phi,
                                         const FilterPair& f, const double* __restrict__ taps, const double (&sreg)[NS],
                                         double2 (&o)[R]) {
  constexpr bool ST = (L > kUniformTapsMaxDwt) && !QMF && !DIRECT;
  constexpr int HL = L / 2;
#pragma unroll
  for (int r = 0; r < R; r++) { o[r].x = 0.0; o[r].y = 0.0; }
  double ts[L], tw[L];
  // children ascending (the reference adds the contributions to one output in ascending child index):
  // child i feeds output pair r with tap pair q = r + HL - 1 - i
#pragma unroll
  for (int i = 0; i < R + HL - 1; ++i) {
    if (ST) {
      const int qn = HL - 1 - i;   // the new (smallest so far) tap pair when i < HL
      if (qn >= 0) {
        ts[2 * qn] = taps[2 * qn];
        ts[2 * qn + 1] = taps[2 * qn + 1];
        tw[2 * qn] = taps[JWC_MAX_TAPS + 2 * qn];
        tw[2 * qn + 1] = taps[JWC_MAX_TAPS + 2 * qn + 1];
      }
    }
    const double cl = plo[i], ch = phi[i];
#pragma unroll
    for (int r = 0; r < R; r++) {
      const int q = r + HL - 1 - i;
      if (q >= 0 && q < HL) {
        if (QMF) {
          o[r].x = fma(ch, sreg[(L - 1 - 2 * q) % NS], fma(cl, sreg[(2 * q) % NS], o[r].x));
          o[r].y = fma(ch, -sreg[(L - 2 - 2 * q) % NS], fma(cl, sreg[(2 * q + 1) % NS], o[r].y));
        } else {
          o[r].x = fma(ch, ST ? tw[2 * q] : f.f1[2 * q], fma(cl, ST ? ts[2 * q] : f.f0[2 * q], o[r].x));
          o[r].y = fma(ch, ST ? tw[2 * q + 1] : f.f1[2 * q + 1], fma(cl, ST ? ts[2 * q + 1] : f.f0[2 * q + 1], o[r].y));
        }
      }
    }
  }
}

__device__ __forceinline__ int inv_halo(int L, int jj) {
  int h = 0;
  for (int q = 1; q <= jj; q++) {
    h = (h + 1) / 2 + (L / 2 - 1);
    h += h & 1;
  }
  return h;
}

template <int L, int R, bool TREE, bool QMF, int NS>
__device__ __forceinline__ void syn_level(double* smem, const FilterPair& f, const double (&sreg)[NS], int oT, int oin,
                                          int oout, int st_in, int st_out, int np, int off, int parents, int tid, int nt) {
  const int nb = (np + R - 1) / R;
  const int items = parents * nb;
  // tap source: constant bank in the packet-tree kernel; the pyramid (FWT) instantiation spills with it, so it keeps
  // the shared-memory copy
  // tap source: constant bank for the packet tree (L <= 20); the pyramid (FWT) instantiation needs 128 registers
  // with it and measured slower (db8 inverse 5.9 ms vs 4.9 ms), so it keeps the shared-memory copy
  const double* ctaps = (TREE && L <= 20) ? const_taps(f) : (smem + oT);
  int zq;
  asm volatile("mov.u32 %0, 0;" : "=r"(zq));   // opaque 0: keeps the pyramid instantiation on the same code shape as the tree one
#pragma unroll 1
  for (int w = tid; w < items; w += nt) {
    const int q = TREE ? (w / nb) : zq;
    const int u0 = (w - q * nb) * R;
    const int clo = oin + (2 * q) * st_in + off + u0;
    double2 o[R];
    syn_item<L, R, QMF, NS, kSynDirect && !TREE && (L > kUniformTapsMaxDwt) && (L <= 20)>(smem + clo, smem + clo + st_in, f, ctaps, sreg, o);
    double2* dst = reinterpret_cast<double2*>(smem + oout + q * st_out) + u0;
    if (u0 + R <= np) {
#pragma unroll
      for (int r = 0; r < R; r++) dst[r] = o[r];
    } else {
#pragma unroll
      for (int r = 0; r < R; r++)
        if (u0 + r < np) dst[r] = o[r];
    }
  }
}

template <int L, int RMAX, bool TREE, bool QMF>
__global__ void __launch_bounds__(256, (TREE ? (L > 6 ? 2 : 3) : ((L >= 8 && L <= 20) ? 2 : 3))) dwt_inv_pass_kernel(const __grid_constant__ DwtPassArgs a,
                                                              const __grid_constant__ FilterPair f) {
  constexpr int NS = QMF ? L : 1;
  double sreg[NS];
  if (QMF) {
    const double* ct = const_taps(f);
#pragma unroll
    for (int m = 0; m < NS; m++) sreg[m] = ct[m];
  } else {
    sreg[0] = 0.0;
  }
  extern __shared__ __align__(128) double smem[];
  __shared__ int s_hl[16];   // halo table (dynamic indexing of a kernel-parameter array would go through local memory)
  const int tid = threadIdx.x, nt = blockDim.x;
  if (tid < 16) s_hl[tid] = a.hl[tid];
  const int oT = 2 * a.cap;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + oT + 2 * JWC_MAX_TAPS);
  if (L > 20 || (L > kUniformTapsMaxDwt && !TREE && !QMF)) {   // the only instantiations that read the shared-memory copy
    for (int t = tid; t < JWC_MAX_TAPS; t += nt) {
      smem[oT + t] = f.f0[t];
      smem[oT + JWC_MAX_TAPS + t] = f.f1[t];
    }
  }
  const unsigned bid = blockIdx.x;
  const int ti = (int)(bid & (unsigned)(a.tiles - 1));
  const int p = (int)((bid >> a.log_tiles) & (unsigned)(a.nodes - 1));
  const int64_t b = (int64_t)(bid >> (a.log_tiles + a.l0 * (a.nodes > 1 ? 1 : 0)));
  const int tlen = (int)((a.h < a.T) ? a.h : a.T);
  const int64_t a0 = (int64_t)ti * tlen;
  const bool bulk = (a.mode == DWT_BULK);
  const double* in_b = a.in + b * a.in_sig;

  // split-device mode: END of halo slot `slot` (the samples before position 0 of that array, right-aligned)
  auto ihalo = [&](int slot) -> const double* { return a.halo ? a.halo + (int64_t)(slot + 1) * a.halo_stride : nullptr; };
  // geometry of depth jj: children arrays of length len = (tlen >> jj) + HLj, node length hn = h >> jj
  auto len_of = [&](int jj) { return (tlen >> jj) + s_hl[jj]; };
  auto stride_of = [&](int jj) { const int l = len_of(jj); return l + (l & 1) + 2 * kDwtR; };

  // upfront: all k detail tiles are requested before the first level runs.  D_{l0+jj-1} lands in the high-pass slot of
  // the buffer level jj writes to, exactly where the one-level-ahead prefetch puts it; tiles of the same buffer nest
  // without overlap ([stride_j, stride_j + len_j) lies below stride_{j-2} when T >> k >= L, checked by the host), so the
  // waits of the deeper levels overlap one DRAM round trip instead of paying one each.  One mbarrier per level, phase 0.
  const bool upfront = bulk && !TREE && a.upfront != 0;
  if (bulk && tid == 0) {
    const int nb = upfront ? a.k : 2;
    for (int q = 0; q < nb; q++) ptx::mbar_init(&bars[q], 1);
    ptx::fence_mbar_init();
    if (!TREE) {
      // the first two tiles (A_{l0+k}, D_{l0+k}) are requested before the block barrier below: thread 0 has just
      // initialised the mbarrier itself, and the halo comes straight from the parameter bank (s_hl is not visible yet)
      const int jj = a.k, hk = a.hl[a.k];
      const int len = (tlen >> jj) + hk;
      const int st = len + (len & 1) + 2 * kDwtR;
      const int64_t hn = a.h >> jj, start = (a0 >> jj) - hk;
      ptx::mbar_expect_tx(&bars[0], 2u * (uint32_t)len * 8u);
      bulk_load_circ(smem, a.ain + b * a.ain_sig, start, len, hn, &bars[0], nullptr, ihalo(0));
      bulk_load_circ(smem + st, in_b + (a.N >> (a.l0 + jj)), start, len, hn, &bars[0], nullptr, ihalo(1));
    }
  }
  __syncthreads();   // mbarriers, s_hl and the tap copy are visible
  // ---- prologue: the depth-k set into buffer 0 ------------------------------------------------------------------------------
  {
    const int jj = a.k, len = len_of(jj), st = stride_of(jj);
    const int64_t hn = a.h >> jj, start = (a0 >> jj) - s_hl[jj];
    if (TREE) {
      const int leaves = 1 << jj;
      const double* src0 = in_b + (int64_t)p * a.h;       // leaf c of this node at + c * hn
      if (bulk) {
        if (tid == 0) ptx::mbar_expect_tx(&bars[0], (uint32_t)(leaves * len) * 8u);
        __syncthreads();
        if (tid < 32)
          for (int c = tid; c < leaves; c += 32)
            bulk_load_circ(smem + c * st, src0 + c * hn, start, len, hn, &bars[0], nullptr, ihalo(p * leaves + c));
      } else {
        for (int c = 0; c < leaves; c++)
          scalar_load_circ(smem + c * st, src0 + c * hn, start, len, hn, tid, nt, nullptr, ihalo(p * leaves + c));
      }
    } else {
      const double* asrc = a.ain + b * a.ain_sig;                 // A_{l0+k}
      const double* dsrc = in_b + (a.N >> (a.l0 + jj));           // D_{l0+k}
      if (bulk) {   // (A_{l0+k} and D_{l0+k} are on their way since the top of the kernel)
        if (upfront && (tid & 31) == 0) {
          // D_{l0+j2} is read by iteration u2 = k - j2 from buffer u2 & 1.  Lane 0 of warp (u2 mod warps) requests it:
          // the requests of the levels leave in parallel instead of queueing behind thread 0's address arithmetic
          const int nw = nt >> 5;
          for (int u2 = (tid >> 5) == 0 ? nw : (tid >> 5); u2 < a.k; u2 += nw) {
            const int j2 = a.k - u2;
            const int len2 = len_of(j2);
            const int64_t hn2 = a.h >> j2, start2 = (a0 >> j2) - s_hl[j2];
            ptx::mbar_expect_tx(&bars[u2], (uint32_t)len2 * 8u);
            bulk_load_circ(smem + (u2 & 1) * a.cap + stride_of(j2), in_b + (a.N >> (a.l0 + j2)), start2, len2, hn2,
                           &bars[u2], nullptr, ihalo(1 + a.k - j2));
          }
        }
      } else {
        scalar_load_circ(smem, asrc, start, len, hn, tid, nt, nullptr, ihalo(0));
        scalar_load_circ(smem + st, dsrc, start, len, hn, tid, nt, nullptr, ihalo(1));
      }
    }
  }
  if (bulk && a.pf_dist > 0 && blockIdx.x + (unsigned)a.pf_dist < a.nblocks) {   // one wave ahead into L2
    const unsigned nb = blockIdx.x + (unsigned)a.pf_dist;
    const int ti2 = (int)(nb & (unsigned)(a.tiles - 1));
    const int p2 = (int)((nb >> a.log_tiles) & (unsigned)(a.nodes - 1));
    const int64_t b2 = (int64_t)(nb >> (a.log_tiles + a.l0 * (a.nodes > 1 ? 1 : 0)));
    const int64_t a2 = (int64_t)ti2 * tlen;
    if (TREE) {
      const int own = tlen >> a.k;
      if (tid < (1 << a.k) && own >= 2)
        ptx::bulk_prefetch_l2(a.in + b2 * a.in_sig + (int64_t)p2 * a.h + (int64_t)tid * (a.h >> a.k) + (a2 >> a.k), (uint32_t)own * 8u);
    } else if (tid <= a.k) {
      const int jj = tid == 0 ? a.k : tid;   // thread 0: A_{l0+k}; thread t: D_{l0+t}
      const int own = tlen >> jj;
      const double* src = tid == 0 ? (a.ain + b2 * a.ain_sig) : (a.in + b2 * a.in_sig + (a.N >> (a.l0 + jj)));
      if (own >= 2) ptx::bulk_prefetch_l2(src + (a2 >> jj), (uint32_t)own * 8u);
    }
  }
  // ---- k synthesis levels: depth jj (buffer u&1) -> depth jj-1 (buffer (u+1)&1) ---------------------------------------
  for (int jj = a.k, u = 0; jj >= 1; --jj, ++u) {
    const int oin = (u & 1) * a.cap, oout = ((u + 1) & 1) * a.cap;
    const int st_in = stride_of(jj), st_out = stride_of(jj - 1);
    if (!TREE && jj > 1 && !upfront) {   // prefetch D_{l0+jj-1} into the high-pass slot of the output buffer
      const int len = len_of(jj - 1);
      const int64_t hn = a.h >> (jj - 1), start = (a0 >> (jj - 1)) - s_hl[jj - 1];
      const double* dsrc = in_b + (a.N >> (a.l0 + jj - 1));
      if (bulk) {
        if (tid == 0) {
          ptx::mbar_expect_tx(&bars[(u + 1) & 1], (uint32_t)len * 8u);
          bulk_load_circ(smem + oout + st_out, dsrc, start, len, hn, &bars[(u + 1) & 1], nullptr, ihalo(1 + a.k - (jj - 1)));
        }
      } else {
        scalar_load_circ(smem + oout + st_out, dsrc, start, len, hn, tid, nt, nullptr, ihalo(1 + a.k - (jj - 1)));
      }
    }
    if (bulk && (!TREE || u == 0)) {
      const uint32_t par = upfront ? 0u : (uint32_t)((u >> 1) & 1);
      uint64_t* lvl_bar = &bars[upfront ? u : (u & 1)];
      if (a.top_barrier == 1) {   // round-1 form: one sleeper on the mbarrier, the rest on a block barrier
        if (tid == 0) ptx::mbar_wait(lvl_bar, par);
        __syncthreads();
      }
      // every thread takes its own acquire on the TMA-written tiles; the block barrier that ends the previous level
      // has already ordered the generic-proxy traffic, so the top of a level needs no second one
      if (a.top_barrier >= 16) ptx::mbar_wait_hint(lvl_bar, par, (uint32_t)a.top_barrier);   // experiment: suspend-time hint in ns
      else ptx::mbar_wait(lvl_bar, par);
    } else if (u == 0 || a.top_barrier == 1) {
      __syncthreads();   // scalar prologue loads (all threads) -> visible
    }
    const int hl_out = s_hl[jj - 1], hl_in = s_hl[jj];
    const int np = (hl_out >> 1) + (tlen >> jj);          // output pairs per parent
    const int off = hl_in - (hl_out >> 1) - (L / 2 - 1);   // first child index read by pair 0
    const int parents = TREE ? (1 << (jj - 1)) : 1;
    const int Rsel = dwt_pick_r(L, np, parents, nt);
    switch (Rsel) {
      case 7:
        if constexpr (RMAX >= 7) {
          syn_level<L, 7, TREE, QMF, NS>(smem, f, sreg, oT, oin, oout, st_in, st_out, np, off, parents, tid, nt);
          break;
        }
      case 5:
        if constexpr (RMAX == 5) {
          syn_level<L, 5, TREE, QMF, NS>(smem, f, sreg, oT, oin, oout, st_in, st_out, np, off, parents, tid, nt);
          break;
        }
      case 3:
        syn_level<L, 3, TREE, QMF, NS>(smem, f, sreg, oT, oin, oout, st_in, st_out, np, off, parents, tid, nt);
        break;
      default:
        syn_level<L, 1, TREE, QMF, NS>(smem, f, sreg, oT, oin, oout, st_in, st_out, np, off, parents, tid, nt);
    }

    if (bulk && jj == 1) ptx::fence_proxy_async();
    __syncthreads();
  }
  // ---- the depth-0 tile leaves ---------------------------------------------------------------------------------------------
  const double* res = smem + (a.k & 1) * a.cap;
  double* dst = a.out + b * a.out_sig + (int64_t)p * a.h + a0;
  if (bulk) {
    if (tid == 0) {
      ptx::bulk_s2g(dst, res, (uint32_t)tlen * 8u);
      ptx::bulk_commit();
      ptx::bulk_wait_read<0>();
    }
  } else {
    for (int e = tid; e < tlen; e += nt) dst[e] = res[e];
  }
}

template <int L, int RMAX, bool TREE, bool INV, bool QMF>
int launch_dwt_pass_r(jwc_ctx* ctx, cudaStream_t st, const DwtPassArgs& a, const FilterPair& f, int threads, size_t smem,
                      int64_t nblocks) {
  if (INV) {
    auto kern = dwt_inv_pass_kernel<L, RMAX, TREE, QMF>;
    JWC_CUDA_CHECK(allow_max_dynamic_smem(kern));
    kern<<<(unsigned)nblocks, threads, smem, st>>>(a, f);
  } else {
    auto kern = dwt_fwd_pass_kernel<L, RMAX, TREE, QMF>;
    JWC_CUDA_CHECK(allow_max_dynamic_smem(kern));
    kern<<<(unsigned)nblocks, threads, smem, st>>>(a, f);
  }
  count_launch(ctx);
  JWC_CUDA_CHECK(cudaGetLastError());
  return JWC_OK;
}

template <int L, bool TREE, bool INV, bool QMF>
int launch_dwt_pass(jwc_ctx* ctx, cudaStream_t st, const DwtPassArgs& a, const FilterPair& f, int threads, size_t smem,
                    int64_t nblocks) {
  return launch_dwt_pass_r<L, dwt_rmax(L), TREE, INV, QMF>(ctx, st, a, f, threads, smem, nblocks);
}

// exact quadrature-mirror relation of the reference's orthogonal wavelets (Wavelet.java:109-113)
bool is_qmf(const FilterPair& f, int L) {
  for (int i = 0; i < L; i++) {
    const double e = (i % 2 == 0) ? f.f0[L - 1 - i] : -f.f0[L - 1 - i];
    if (!(f.f1[i] == e)) return false;
  }
  return true;
}

template <bool TREE, bool INV>
int dispatch_dwt_pass(jwc_ctx* ctx, cudaStream_t st, const DwtPassArgs& a, const FilterPair& f, int L, int threads,
                      size_t smem, int64_t nblocks) {
  // forward, and the pyramid inverse (its non-QMF instantiation reads the taps from shared memory with broadcast LDS
  // and runs reg,reg,reg DFMAs; the QMF one keeps the L distinct tap values in uniform registers).  The packet-tree
  // inverse already has uniform-register taps without it.
  if constexpr (!INV || !TREE) if (L >= 12 && L <= 20 && ctx->tune.dwt_qmf >= 0 && is_qmf(f, L)) {
    switch (L) {
#define JWC_QCASE(LL) case LL: return launch_dwt_pass<LL, TREE, INV, true>(ctx, st, a, f, threads, smem, nblocks);
      JWC_QMF_L(JWC_QCASE)
#undef JWC_QCASE
      default: break;
    }
  }
  switch (L) {
#define JWC_CASE(LL) case LL: return launch_dwt_pass<LL, TREE, INV, false>(ctx, st, a, f, threads, smem, nblocks);
    JWC_ALL_L(JWC_CASE)
#undef JWC_CASE
    default: return JWC_ERR_UNSUPPORTED;
  }
}

int steps_forward(int64_t n, int levels) {
  int steps = 0;
  for (int64_t h = n; h >= 2 && steps < levels; h >>= 1) steps++;
  return steps;
}

int dwt_prefetch_distance(jwc_ctx* ctx, const DeviceSlot& dev, size_t smem, int threads) {
  (void)smem; (void)threads;
  if (ctx->tune.l2_prefetch < 0) return 0;
  if (ctx->tune.l2_prefetch > 0) return ctx->tune.l2_prefetch;
  return 2 * dev.sm_count;   // measured on B200 (FWT Haar): 2 CTAs per SM ahead 2.62 ms, a wave ahead 2.72, off 2.95
}

DwtPlan make_plan(jwc_ctx* ctx, const DeviceSlot& dev, const void* p0, const void* p1, int64_t n, int steps, int L,
                  bool tree, bool inverse, int64_t ld, int group_cap = 0) {
  DwtPlanInput pin{};
  pin.n = n; pin.levels = steps; pin.L = L; pin.tree = tree; pin.inverse = inverse;
  pin.aligned16 = ((reinterpret_cast<uintptr_t>(p0) | reinterpret_cast<uintptr_t>(p1)) & 15) == 0 && (ld & 1) == 0;
  pin.smem_budget = ctx->tune.dwt_smem > 0 ? ctx->tune.dwt_smem : 45000;
  if (pin.smem_budget > dev.max_smem_optin) pin.smem_budget = dev.max_smem_optin;
  pin.tile_override = ctx->tune.dwt_tile; pin.group_override = ctx->tune.dwt_group;
  // pyramid inverse: every level waits for its D tile (prefetched only one level ahead), so short passes win
  // (measured, Haar 2^20: k <= 3 gives 3.10 ms, the model's k = 5 gives 3.38 ms)
  if (inverse && !tree && pin.group_override <= 0) pin.group_override = 3;
  pin.threads_override = ctx->tune.dwt_threads;
  pin.k0_override = ctx->tune.dwt_k0;
  pin.fixed_override = ctx->tune.dwt_fixed;
  if (group_cap > 0) pin.group_override = pin.group_override > 0 ? std::min(pin.group_override, group_cap) : group_cap;
  return dwt_plan(pin, steps);
}

}  // namespace

int fast_dwt_forward(jwc_ctx* ctx, const DeviceSlot& dev, cudaStream_t st, const double* d_in, double* d_out,
                     int64_t batch, int64_t n, int levels, const FilterPair& f, int L, bool tree, int64_t ld) {
  if (ld <= 0) ld = n;
  if (L < 2 || L > 40 || (L & 1)) return JWC_ERR_UNSUPPORTED;
  if (n >= ((int64_t)1 << 31)) return JWC_ERR_UNSUPPORTED;
  const int steps = steps_forward(n, levels);
  if (steps == 0) return JWC_ERR_UNSUPPORTED;   // plain copy: the generic path handles it
  if (tree && whole_dwt_levels(ctx, n, steps, L, true) == steps) {   // short packets transforms: all blocks in place
    const int rc = whole_dwt(ctx, st, d_in, d_out, batch, n, steps, f, L, ld, false, nullptr, 0, true);
    if (rc != JWC_ERR_UNSUPPORTED) return rc;
  }
  if (!tree) {   // 512 < n <= 4096: the whole signal in shared memory, levels in place (jwc_dwt_whole.cu)
    // (forward: only when it takes every level -- for a deep transform the first tile pass + the warp tail measured
    // faster than this kernel + the tail, 1.73 vs 1.83 ms on 131 072 rows of 4096; the inverse is the other way round)
    if (whole_dwt_levels(ctx, n, steps, L) == steps) return whole_dwt(ctx, st, d_in, d_out, batch, n, steps, f, L, ld, false);
  }
  // short signals: levels below kDwtTailLen samples go to the warp-per-signal tail kernel (jwc_dwt_tail.cu)
  const int lt = (!tree && ctx->tune.dwt_tail >= 0) ? dwt_tail_start(n, steps) : -1;
  if (lt == 0) return dwt_tail_forward(ctx, st, d_in, ld, d_out, ld, (int)n, steps, batch, f, L);
  const int plan_steps = lt > 0 ? lt : steps;
  const DwtPlan plan = make_plan(ctx, dev, d_in, d_out, n, plan_steps, L, tree, false, ld);
  if (!plan.ok) return JWC_ERR_UNSUPPORTED;
  debug_dwt_plan("forward", plan, n, levels, L, tree);
  Scratch ws(ctx, dev, st);
  const int npass = (int)plan.passes.size();
  double* tail_a = nullptr;
  if (lt > 0) {
    tail_a = ws.get((size_t)batch * (n >> lt));
    if (!tail_a) { set_error("scratch allocation failed"); return JWC_ERR_NOMEM; }
  }
  double* tmp[2] = {nullptr, nullptr};
  if (npass >= 2) {
    if (tree) {
      tmp[0] = ws.get((size_t)batch * n);
      tmp[1] = tmp[0];
    } else {
      tmp[0] = ws.get((size_t)batch * (n >> plan.passes[0].k));
      tmp[1] = (npass >= 3) ? ws.get((size_t)batch * (n >> (plan.passes[0].k + plan.passes[1].k))) : tmp[0];
    }
    if (!tmp[0] || !tmp[1]) { set_error("scratch allocation failed"); return JWC_ERR_NOMEM; }
  }
  const double* src = d_in;
  int64_t src_sig = ld;
  for (int pi = 0; pi < npass; pi++) {
    const DwtPass& p = plan.passes[pi];
    const bool lastp = (pi == npass - 1);
    DwtPassArgs a{};
    a.in = src; a.in_sig = src_sig;
    a.N = n; a.h = n >> p.l0; a.l0 = p.l0; a.k = p.k; a.T = p.T; a.cap = p.cap; a.mode = p.mode;
    a.tiles = (int)((a.h + p.T - 1) / p.T);
    if (tree) {
      a.nodes = 1 << p.l0;
      // whole-array ping-pong so that the last pass lands in d_out
      double* dst = (((npass - 1 - pi) & 1) == 0) ? d_out : tmp[0];
      a.out = dst; a.out_sig = (dst == d_out) ? ld : n;
    } else {
      a.nodes = 1;
      a.out = d_out; a.out_sig = ld;                                 // D's at their final place
      if (lastp && lt > 0) { a.aout = tail_a; a.aout_sig = n >> lt; }
      else if (lastp) { a.aout = d_out; a.aout_sig = ld; }
      else { a.aout = tmp[pi & 1]; a.aout_sig = n >> (p.l0 + p.k); }
    }
    const int64_t nblocks = (int64_t)a.tiles * a.nodes * batch;
    if (nblocks > 0x7fffffffLL) return JWC_ERR_UNSUPPORTED;
    a.nblocks = (unsigned)nblocks;
    a.log_tiles = 0;
    while ((1 << a.log_tiles) < a.tiles) a.log_tiles++;   // tiles = h / T, both powers of two
    a.pf_dist = (p.mode == DWT_BULK) ? dwt_prefetch_distance(ctx, dev, p.smem, p.threads) : 0;
    int rc = tree ? dispatch_dwt_pass<true, false>(ctx, st, a, f, L, p.threads, p.smem, nblocks)
                  : dispatch_dwt_pass<false, false>(ctx, st, a, f, L, p.threads, p.smem, nblocks);
    if (rc != JWC_OK) return rc;
    if (tree) { src = a.out; src_sig = a.out_sig; }
    else { src = a.aout; src_sig = a.aout_sig; }
  }
  if (lt > 0) return dwt_tail_forward(ctx, st, tail_a, n >> lt, d_out, ld, (int)(n >> lt), steps - lt, batch, f, L);
  return JWC_OK;
}

int fast_dwt_inverse(jwc_ctx* ctx, const DeviceSlot& dev, cudaStream_t st, const double* d_in, double* d_out,
                     int64_t batch, int64_t n, int levels, const FilterPair& f, int L, bool tree, int64_t ld) {
  if (ld <= 0) ld = n;
  if (L < 2 || L > 40 || (L & 1)) return JWC_ERR_UNSUPPORTED;
  if (n >= ((int64_t)1 << 31)) return JWC_ERR_UNSUPPORTED;
  const int steps = steps_forward(n, levels);   // the reverse loops undo exactly the forward's steps
  if (steps == 0) return JWC_ERR_UNSUPPORTED;
  if (tree && whole_dwt_levels(ctx, n, steps, L, true) == steps) {
    const int rc = whole_dwt(ctx, st, d_in, d_out, batch, n, steps, f, L, ld, true, nullptr, 0, true);
    if (rc != JWC_ERR_UNSUPPORTED) return rc;
  }
  if (!tree) {
    const int top = whole_dwt_levels(ctx, n, steps, L);
    if (top == steps) return whole_dwt(ctx, st, d_in, d_out, batch, n, steps, f, L, ld, true);
    if (top > 0) {
      // deep end first (warp per signal) into scratch, then the big levels with that approximation as their head
      Scratch ws(ctx, dev, st);
      const int64_t h0 = n >> top;
      double* a_top = ws.get((size_t)batch * h0);
      if (!a_top) { set_error("scratch allocation failed"); return JWC_ERR_NOMEM; }
      const int rc = dwt_tail_inverse(ctx, st, d_in, ld, a_top, h0, (int)h0, steps - top, batch, f, L);
      if (rc != JWC_OK) return rc;
      return whole_dwt(ctx, st, d_in, d_out, batch, n, top, f, L, ld, true, a_top, h0);
    }
  }
  const int lt = (!tree && ctx->tune.dwt_tail >= 0) ? dwt_tail_start(n, steps) : -1;
  if (lt == 0) return dwt_tail_inverse(ctx, st, d_in, ld, d_out, ld, (int)n, steps, batch, f, L);
  const int plan_steps = lt > 0 ? lt : steps;
  const DwtPlan plan = make_plan(ctx, dev, d_in, d_out, n, plan_steps, L, tree, true, ld);
  if (!plan.ok) return JWC_ERR_UNSUPPORTED;
  debug_dwt_plan("inverse", plan, n, levels, L, tree);
  Scratch ws(ctx, dev, st);
  const int npass = (int)plan.passes.size();
  double* tail_a = nullptr;
  if (lt > 0) {
    // the tail rebuilds A at level lt from the deepest approximation and the detail blocks below it
    tail_a = ws.get((size_t)batch * (n >> lt));
    if (!tail_a) { set_error("scratch allocation failed"); return JWC_ERR_NOMEM; }
    const int rc = dwt_tail_inverse(ctx, st, d_in, ld, tail_a, n >> lt, (int)(n >> lt), steps - lt, batch, f, L);
    if (rc != JWC_OK) return rc;
  }
  double* tmp[2] = {nullptr, nullptr};
  if (npass >= 2) {
    if (tree) {
      tmp[0] = ws.get((size_t)batch * n);
      tmp[1] = tmp[0];
    } else {
      // A buffers: the result of pass pi (depth l0_pi) has n >> l0_pi samples; the largest intermediate is pass 1's
      tmp[0] = ws.get((size_t)batch * (n >> plan.passes[1].l0));
      tmp[1] = (npass >= 3) ? ws.get((size_t)batch * (n >> plan.passes[2].l0)) : tmp[0];
    }
    if (!tmp[0] || !tmp[1]) { set_error("scratch allocation failed"); return JWC_ERR_NOMEM; }
  }
  const double* asrc = d_in;     // FWT: A_{steps} is the prefix of the coefficient array; WPT: the whole array
  int64_t asrc_sig = ld;
  if (lt > 0) { asrc = tail_a; asrc_sig = n >> lt; }
  for (int pi = npass - 1, step = 0; pi >= 0; --pi, ++step) {   // deepest pass first
    const DwtPass& p = plan.passes[pi];
    DwtPassArgs a{};
    a.N = n; a.h = n >> p.l0; a.l0 = p.l0; a.k = p.k; a.T = p.T; a.cap = p.cap; a.mode = p.mode;
    a.tiles = (int)((a.h + p.T - 1) / p.T);
    for (int jj = 0; jj < 16; jj++) a.hl[jj] = (int)dwt_inv_halo(L, jj);
    if (tree) {
      a.nodes = 1 << p.l0;
      a.in = asrc; a.in_sig = asrc_sig;
      double* dst = (((pi) & 1) == 0) ? d_out : tmp[0];   // pass 0 (executed last) lands in d_out
      a.out = dst; a.out_sig = (dst == d_out) ? ld : n;
    } else {
      a.nodes = 1;
      a.in = d_in; a.in_sig = ld;                // D blocks always come from the coefficient array
      a.ain = asrc; a.ain_sig = asrc_sig;
      if (pi == 0) { a.out = d_out; a.out_sig = ld; }
      else { a.out = tmp[(pi - 1) & 1]; a.out_sig = n >> p.l0; }
    }
    const int64_t nblocks = (int64_t)a.tiles * a.nodes * batch;
    if (nblocks > 0x7fffffffLL) return JWC_ERR_UNSUPPORTED;
    a.nblocks = (unsigned)nblocks;
    a.log_tiles = 0;
    while ((1 << a.log_tiles) < a.tiles) a.log_tiles++;   // tiles = h / T, both powers of two
    a.pf_dist = (p.mode == DWT_BULK) ? dwt_prefetch_distance(ctx, dev, p.smem, p.threads) : 0;
    if (ctx->tune.pf_inv != 0) a.pf_dist = (p.mode == DWT_BULK && ctx->tune.pf_inv > 0) ? ctx->tune.pf_inv : 0;
    a.top_barrier = ctx->tune.top_barrier;
    a.upfront = 0;
    // (k <= 15: one mbarrier per level in 16 slots and the 16-entry halo table; whole warps: lane 0 of warp u requests tile u)
    if (!tree && p.mode == DWT_BULK && ctx->tune.dwt_upfront > 0 && p.k >= 2 && p.k <= 15 && (p.threads & 31) == 0) {
      // the detail tiles of one buffer must nest: slot of D_j = [stride_j, stride_j + len_j) below the slot of D_{j-2}
      const int64_t tl = std::min<int64_t>(a.h, p.T);
      bool ok = true;
      for (int j = 3; j <= p.k && ok; j++) {
        const int64_t lj = dwt_inv_len(L, j, tl), l2 = dwt_inv_len(L, j - 2, tl);
        ok = dwt_node_stride(l2) >= dwt_node_stride(lj) + lj;
      }
      a.upfront = ok ? 1 : 0;
    }
    int rc = JWC_ERR_UNSUPPORTED;
    if (!tree)   // long signals: the tiled in-place kernel (jwc_dwt_whole.cu) takes passes of up to 3 levels
      rc = tile_dwt_inverse_pass(ctx, st, a.ain, a.ain_sig, a.in, a.in_sig, a.out, a.out_sig, n, p.l0, p.k, batch, f, L);
    if (rc == JWC_ERR_UNSUPPORTED)
      rc = tree ? dispatch_dwt_pass<true, true>(ctx, st, a, f, L, p.threads, p.smem, nblocks)
                : dispatch_dwt_pass<false, true>(ctx, st, a, f, L, p.threads, p.smem, nblocks);
    if (rc != JWC_OK) return rc;
    asrc = a.out; asrc_sig = a.out_sig;
  }
  return JWC_OK;
}

// =========================================================================================================================
// One long series split over the context's devices (FWT / WPT part).
//
// Device slot p holds samples [p*len, (p+1)*len) of the series, len = n / P (n and P powers of two).  Every analysis
// level halves the local part: after l levels slot p owns the coefficients whose time support starts in its chunk,
// (len >> l) per node, so the data never moves between devices -- only the halo does.  A fused pass of k levels reads
// (L-2)(2^k-1) samples past the end of each local node part (forward) or at most L-2 coefficients before the start of
// each child part (inverse): those come from the ring neighbour with cudaMemcpyPeerAsync (one copy per node / child
// array), and the tile kernels take them from the small halo buffer instead of wrapping periodically.  Passes are
// ordered across devices with events (a pass waits for both neighbours' previous pass); no collective, no host sync
// inside.
//
// Result layout ("local layout"): chunk p of the output is the transform's own layout of a signal of length len, filled
// with the coefficients slot p owns:
//   WPT, J levels:  [leaf 0 part | leaf 1 part | ... ], part c = leaf c of the reference, positions [p*len/2^J, (p+1)*len/2^J)
//   FWT:            [T_p | D_ls part | ... | D_1 part], D_l part = D_l[p*len/2^l, (p+1)*len/2^l); the first ls levels
//                   run split (ls = dwt_split_levels); the remaining pyramid works on A_ls, n/2^ls samples -- the
//                   "small remainder": it is gathered on slot 0, transformed there, and its result array
//                   [A_J | D_J .. D_{ls+1}] is cut into P equal contiguous pieces T_p.
// Concatenating the parts of one band over p gives that band of the reference's array, bit for bit: the arithmetic per
// coefficient is the same tile kernel code on the same samples.
// =========================================================================================================================
int dwt_split_levels(int64_t n, int P, int steps, bool tree) {
  (void)tree;
  const int64_t len = n / P;
  int ls = 0;
  while (ls < steps && (len >> ls) >= 2048) ls++;   // a level runs split while its input part has >= 2048 samples
  return ls;
}

namespace {

struct SplitDev {
  const DeviceSlot* dev = nullptr;
  std::vector<cudaEvent_t> done;   // done[i]: pass i (in execution order) finished on this device
};

int split_fail(const char* what, cudaError_t e) {
  set_error("split transform: %s failed: %s", what, cudaGetErrorString(e));
  return JWC_ERR_CUDA;
}

}  // namespace

int split_dwt(jwc_ctx* ctx, bool inverse, bool tree, const double* const* d_in, double* const* d_out, int64_t n,
              int levels, const FilterPair& f, int L) {
  const int P = (int)ctx->slots.size();
  if (P < 1 || (P & (P - 1)) || n < 2 || (n & (n - 1)) || n % P || n >= ((int64_t)1 << 40)) {
    set_error("split FWT/WPT needs a power-of-two length (got %lld) on a power-of-two number of devices (got %d)",
              (long long)n, P);
    return JWC_ERR_INVALID;
  }
  if (L < 2 || L > 40 || (L & 1)) return JWC_ERR_UNSUPPORTED;
  const int64_t len = n / P;
  if (len >= ((int64_t)1 << 31)) return JWC_ERR_UNSUPPORTED;
  const int steps = steps_forward(n, levels);
  const int ls = dwt_split_levels(n, P, steps, tree);
  if (tree && steps > ls) {
    set_error("split WPT: %d levels would leave packet parts shorter than 1024 samples per device (at most %d here)", steps, ls);
    return JWC_ERR_UNSUPPORTED;
  }
  int prev_dev = 0;
  cudaGetDevice(&prev_dev);
  int rc = JWC_OK;
  cudaError_t e = cudaSuccess;
  std::vector<std::unique_ptr<Scratch>> ws;
  for (int p = 0; p < P; p++) ws.emplace_back(new Scratch(ctx, ctx->slots[p], ctx->slots[p].stream));
  auto stream = [&](int p) { return ctx->slots[p].stream; };
  auto setdev = [&](int p) { return cudaSetDevice(ctx->slots[p].ordinal); };
  auto alloc = [&](int p, int64_t doubles) -> double* {
    setdev(p);
    double* q = ws[p]->get((size_t)doubles);
    if (!q) { set_error("split transform: scratch allocation failed on slot %d", p); rc = JWC_ERR_NOMEM; }
    return q;
  };

  // ---- plan of the split levels (on one chunk) ----------------------------------------------------------------------
  DwtPlan plan;
  if (ls > 0) {
    plan = make_plan(ctx, ctx->slots[0], d_in[0], d_out[0], len, ls, L, tree, inverse, len, 3);   // halo <= 7 (L - 2) per pass
    bool aligned = true;
    for (int p = 0; p < P; p++)
      aligned = aligned && ((reinterpret_cast<uintptr_t>(d_in[p]) | reinterpret_cast<uintptr_t>(d_out[p])) & 15) == 0;
    if (!plan.ok) return JWC_ERR_UNSUPPORTED;
    if (!aligned) for (DwtPass& ps : plan.passes) ps.mode = DWT_SCALAR;
  }
  const int npass = (int)plan.passes.size();
  std::vector<SplitDev> sd(P);
  std::vector<cudaEvent_t> tail_ev(P, nullptr);
  for (int p = 0; p < P && rc == JWC_OK; p++) {
    setdev(p);
    sd[p].dev = &ctx->slots[p];
    sd[p].done.assign(npass, nullptr);
    for (int i = 0; i < npass; i++)
      if ((e = cudaEventCreateWithFlags(&sd[p].done[i], cudaEventDisableTiming)) != cudaSuccess) { rc = split_fail("cudaEventCreate", e); break; }
    if (rc == JWC_OK && (e = cudaEventCreateWithFlags(&tail_ev[p], cudaEventDisableTiming)) != cudaSuccess) rc = split_fail("cudaEventCreate", e);
  }
  // copy of `count` doubles from slot q's memory into slot p's memory, issued on slot p's stream
  auto peer = [&](int p, double* dst, int q, const double* src, int64_t count) {
    if (rc != JWC_OK || count <= 0) return;
    setdev(p);
    e = cudaMemcpyPeerAsync(dst, ctx->slots[p].ordinal, src, ctx->slots[q].ordinal, (size_t)count * sizeof(double), stream(p));
    if (e != cudaSuccess) rc = split_fail("cudaMemcpyPeerAsync", e);
  };
  auto wait_ev = [&](int p, cudaEvent_t ev) {
    if (rc != JWC_OK || !ev) return;
    if ((e = cudaStreamWaitEvent(stream(p), ev, 0)) != cudaSuccess) rc = split_fail("cudaStreamWaitEvent", e);
  };
  auto record = [&](int p, cudaEvent_t ev) {
    if (rc != JWC_OK) return;
    if ((e = cudaEventRecord(ev, stream(p))) != cudaSuccess) rc = split_fail("cudaEventRecord", e);
  };
  auto launch = [&](int p, DwtPassArgs& a, const DwtPass& ps) {
    if (rc != JWC_OK) return;
    setdev(p);
    a.N = len; a.l0 = ps.l0; a.k = ps.k; a.T = ps.T; a.cap = ps.cap; a.mode = ps.mode;
    a.h = len >> ps.l0;
    a.tiles = (int)((a.h + ps.T - 1) / ps.T);
    a.nodes = tree ? (1 << ps.l0) : 1;
    for (int jj = 0; jj < 16; jj++) a.hl[jj] = (int)dwt_inv_halo(L, jj);
    const int64_t nblocks = (int64_t)a.tiles * a.nodes;
    a.nblocks = (unsigned)nblocks;
    a.log_tiles = 0;
    while ((1 << a.log_tiles) < a.tiles) a.log_tiles++;
    a.pf_dist = 0;
    const int r = tree ? (inverse ? dispatch_dwt_pass<true, true>(ctx, stream(p), a, f, L, ps.threads, ps.smem, nblocks)
                                  : dispatch_dwt_pass<true, false>(ctx, stream(p), a, f, L, ps.threads, ps.smem, nblocks))
                       : (inverse ? dispatch_dwt_pass<false, true>(ctx, stream(p), a, f, L, ps.threads, ps.smem, nblocks)
                                  : dispatch_dwt_pass<false, false>(ctx, stream(p), a, f, L, ps.threads, ps.smem, nblocks));
    if (r != JWC_OK) rc = r;
  };
  // ordinary (unsplit) transform of the gathered remainder on slot 0
  auto whole_on_slot0 = [&](const double* src, double* dst, int64_t m, int lv, bool inv) {
    if (rc != JWC_OK) return;
    setdev(0);
    int r = inv ? fast_dwt_inverse(ctx, ctx->slots[0], stream(0), src, dst, 1, m, lv, f, L, tree)
                : fast_dwt_forward(ctx, ctx->slots[0], stream(0), src, dst, 1, m, lv, f, L, tree);
    if (r == JWC_ERR_UNSUPPORTED)
      r = inv ? generic_dwt_inverse(ctx, ctx->slots[0], stream(0), src, dst, 1, m, lv, f, L, tree, false)
              : generic_dwt_forward(ctx, ctx->slots[0], stream(0), src, dst, 1, m, lv, f, L, tree, false);
    if (r != JWC_OK) rc = r;
  };

  if (steps == 0) {   // zero levels: the transform is a copy
    for (int p = 0; p < P; p++) peer(p, d_out[p], p, d_in[p], len);
    for (int p = 0; p < P; p++) { setdev(p); cudaStreamSynchronize(stream(p)); }
    cudaSetDevice(prev_dev);
    return rc;
  }
  const int64_t tail_len = n >> ls;          // the remainder A_ls (FWT; the whole series when ls == 0)
  const int64_t tail_part = tail_len / P;    // = len >> ls
  const int tail_levels = steps - ls;

  if (!inverse) {
    // ================= forward =================
    // per pass and device: source (A_{l0} part / node parts) and destinations
    std::vector<std::vector<const double*>> src(npass + 1, std::vector<const double*>(P, nullptr));
    std::vector<std::vector<double*>> dst(npass, std::vector<double*>(P, nullptr));    // tree: leaves; pyramid: A_{l0+k}
    std::vector<std::vector<double*>> hb(npass, std::vector<double*>(P, nullptr));
    for (int p = 0; p < P && rc == JWC_OK; p++) {
      src[0][p] = d_in[p];
      double* tmp = (tree && npass >= 2) ? alloc(p, len) : nullptr;
      for (int i = 0; i < npass && rc == JWC_OK; i++) {
        const DwtPass& ps = plan.passes[i];
        const int64_t H = (int64_t)(L - 2) * ((1 << ps.k) - 1);
        hb[i][p] = alloc(p, std::max<int64_t>(2, (tree ? ((int64_t)1 << ps.l0) : 1) * H));
        if (tree) dst[i][p] = (((npass - 1 - i) & 1) == 0) ? d_out[p] : tmp;   // whole-array ping-pong, last pass in d_out
        else dst[i][p] = (i == npass - 1 && tail_levels == 0) ? d_out[p] : alloc(p, len >> (ps.l0 + ps.k));
        src[i + 1][p] = dst[i][p];
      }
    }
    for (int i = 0; i < npass && rc == JWC_OK; i++) {
      const DwtPass& ps = plan.passes[i];
      const int64_t h = len >> ps.l0, H = (int64_t)(L - 2) * ((1 << ps.k) - 1);
      const int nodes = tree ? (1 << ps.l0) : 1;
      for (int p = 0; p < P && rc == JWC_OK; p++) {
        setdev(p);
        const int right = (p + 1) % P, left = (p + P - 1) % P;
        if (i > 0) { wait_ev(p, sd[right].done[i - 1]); wait_ev(p, sd[left].done[i - 1]); }
        for (int r = 0; r < nodes; r++) peer(p, hb[i][p] + (int64_t)r * H, right, src[i][right] + (int64_t)r * h, H);
        DwtPassArgs a{};
        a.in = src[i][p]; a.in_sig = len;
        a.halo = hb[i][p]; a.halo_stride = H;
        if (tree) { a.out = dst[i][p]; a.out_sig = len; }
        else { a.out = d_out[p]; a.out_sig = len; a.aout = dst[i][p]; a.aout_sig = len >> (ps.l0 + ps.k); }
        launch(p, a, ps);
        record(p, sd[p].done[i]);
      }
    }
    if (!tree && tail_levels > 0 && rc == JWC_OK) {
      // gather A_ls on slot 0, finish the pyramid there, hand every slot its piece of the result array
      double* tin = alloc(0, tail_len);
      double* tout = alloc(0, tail_len);
      for (int p = 0; p < P && rc == JWC_OK; p++) {
        if (npass > 0) wait_ev(0, sd[p].done[npass - 1]);
        peer(0, tin + (int64_t)p * tail_part, p, src[npass][p], tail_part);
      }
      whole_on_slot0(tin, tout, tail_len, tail_levels, false);
      record(0, tail_ev[0]);
      for (int p = 0; p < P && rc == JWC_OK; p++) {   // every slot fetches its piece on its own stream
        wait_ev(p, tail_ev[0]);
        peer(p, d_out[p], 0, tout + (int64_t)p * tail_part, tail_part);
      }
    }
  } else {
    // ================= inverse =================
    // (1) the remainder: pieces T_p -> slot 0, inverse pyramid there, A_ls parts back to their slots
    std::vector<const double*> asrc(P, nullptr);   // pyramid: A_{deepest split level} part; tree: whole local array
    if (!tree) {
      if (tail_levels > 0) {
        double* tin = alloc(0, tail_len);
        double* tout = alloc(0, tail_len);
        for (int p = 0; p < P && rc == JWC_OK; p++) peer(0, tin + (int64_t)p * tail_part, p, d_in[p], tail_part);
        whole_on_slot0(tin, tout, tail_len, tail_levels, true);
        record(0, tail_ev[0]);
        for (int p = 0; p < P && rc == JWC_OK; p++) {   // every slot fetches its A_ls part on its own stream
          double* ap = (npass == 0) ? d_out[p] : alloc(p, tail_part);
          wait_ev(p, tail_ev[0]);
          peer(p, ap, 0, tout + (int64_t)p * tail_part, tail_part);
          asrc[p] = ap;
          record(p, tail_ev[p]);   // (slot 0 re-records its own event after its copy: later waits see both)
        }
      } else {
        for (int p = 0; p < P; p++) asrc[p] = d_in[p];   // A_ls part is the head of the local layout
      }
    } else {
      for (int p = 0; p < P; p++) asrc[p] = d_in[p];
    }
    // (2) the split passes, deepest first
    std::vector<double*> tmp(P, nullptr);
    if (tree && npass >= 2) for (int p = 0; p < P && rc == JWC_OK; p++) tmp[p] = alloc(p, len);
    for (int i = npass - 1, step = 0; i >= 0 && rc == JWC_OK; --i, ++step) {
      const DwtPass& ps = plan.passes[i];
      const int64_t h = len >> ps.l0, hc = h >> ps.k;
      const int64_t hs = dwt_inv_halo(L, ps.k);       // slot size = left halo of the deepest children
      const int nodes = tree ? (1 << ps.l0) : 1, leaves = 1 << ps.k;
      const int slots = tree ? nodes * leaves : ps.k + 1;
      std::vector<double*> out(P, nullptr), hbuf(P, nullptr);
      for (int p = 0; p < P && rc == JWC_OK; p++) {
        hbuf[p] = alloc(p, std::max<int64_t>(2, (int64_t)slots * hs));
        if (tree) out[p] = ((i & 1) == 0) ? d_out[p] : tmp[p];   // pass 0 (executed last) lands in d_out
        else out[p] = (i == 0) ? d_out[p] : alloc(p, h);
      }
      for (int p = 0; p < P && rc == JWC_OK; p++) {
        setdev(p);
        const int right = (p + 1) % P, left = (p + P - 1) % P;
        if (step > 0) { wait_ev(p, sd[left].done[i + 1]); wait_ev(p, sd[right].done[i + 1]); }
        else if (!tree && tail_levels > 0) wait_ev(p, tail_ev[left]);   // the left neighbour's A_ls part has arrived
        if (tree) {
          for (int r = 0; r < slots; r++)   // child r (node-major): left neighbour's last hs coefficients
            peer(p, hbuf[p] + (int64_t)r * hs, left, asrc[left] + (int64_t)(r + 1) * hc - hs, hs);
        } else {
          peer(p, hbuf[p], left, asrc[left] + hc - hs, hs);                       // slot 0: A_{l0+k}
          for (int jj = ps.k; jj >= 1; --jj) {                                     // slot 1 + (k - jj): D_{l0+jj}
            const int64_t hj = dwt_inv_halo(L, jj), part = h >> jj;
            peer(p, hbuf[p] + (int64_t)(2 + ps.k - jj) * hs - hj, left, d_in[left] + (len >> (ps.l0 + jj)) + part - hj, hj);
          }
        }
        DwtPassArgs a{};
        a.halo = hbuf[p]; a.halo_stride = hs;
        if (tree) { a.in = asrc[p]; a.in_sig = len; a.out = out[p]; a.out_sig = len; }
        else { a.in = d_in[p]; a.in_sig = len; a.ain = asrc[p]; a.ain_sig = hc; a.out = out[p]; a.out_sig = (i == 0) ? len : h; }
        launch(p, a, ps);
        record(p, sd[p].done[i]);
      }
      for (int p = 0; p < P; p++) asrc[p] = out[p];
    }
  }
  // ---- drain every device, release events and workspace ---------------------------------------------------------------
  for (int p = 0; p < P; p++) {
    setdev(p);
    cudaError_t e2 = cudaStreamSynchronize(stream(p));
    if (e2 != cudaSuccess && rc == JWC_OK) rc = split_fail("cudaStreamSynchronize", e2);
    for (cudaEvent_t ev : sd[p].done) if (ev) cudaEventDestroy(ev);
    if (tail_ev[p]) cudaEventDestroy(tail_ev[p]);
  }
  ws.clear();
  cudaSetDevice(prev_dev);
  return rc;
}

}  // namespace jwc
