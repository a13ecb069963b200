// FP64 tensor-core GEMM for sm_100a:  C[M,N] (+)= A[M,K] * B[N,K]^T  with both operands K-major
// (row-major, contraction index contiguous), every dimension a multiple of 128.
//
//  * operands arrive by TMA (cp.async.bulk.tensor.3d, SWIZZLE_128B, box 16 x 128 doubles = 16 KiB) into a
//    4-stage shared-memory ring guarded by full/empty mbarriers; one producer warp, eight consumer warps
//  * consumers run mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4; tcgen05 has no f64 kind), warp tile 64 x 32,
//    fragments fetched with conflict-free 16-byte LDS (k index permuted identically for A and B)
//  * epilogues are fused: plain/transposed store (Cholesky, inverse, W^-1), column sum of squares
//    (posterior variance, value-only path) and the variance-gradient reduction (k.b and sum_j b_j dk_j/dx)
//
// This one kernel is the dense O(n^2)-per-candidate contraction of the path and the
// workhorse of the refit (K4).  Reference arithmetic it replaces: GPy PosteriorExact._raw_predict
// (dtrtrs + sum of squares) and GP.predictive_gradients (K* . woodbury_inv GEMM) as called from
// /root/reference/GPyOpt/models/gpmodel.py:98,136-138.
#pragma once
#include "common.cuh"

namespace gpo {

constexpr int GEMM_BK = 16;                         // k extent of one pipeline stage (16 doubles = 128 B)
constexpr int GEMM_TILE_BYTES = TILE * GEMM_BK * 8;  // 16 KiB: the A operand of one stage (128 rows)

// Two shapes of the same kernel:
//   BN = 128: CTA tile 128 x 128, 8 consumer warps (2 warpgroups) + 1 producer warpgroup, one CTA per SM
//   BN =  64: CTA tile 128 x  64, 4 consumer warps (1 warpgroup)  + 1 producer warpgroup, TWO CTAs per SM.  The two
//             co-resident CTAs are never in step, so one CTA's epilogue / pipeline fill runs under the other's DMMA
//             main loop (with one CTA per SM every epilogue idles the FP64 pipe), and work items are half as long,
//             which shortens the latency-bound launches of the refit.  Price: the A operand is fetched once per 64
//             instead of once per 128 columns (1.5x the L2 -> shared-memory traffic, far below the L2 limit).
template <int BN>
struct GemmShape {
  static constexpr int CW = BN / 16;                           // consumer warps (warp tile 64 x 32)
  static constexpr int THREADS = (CW + 4) * 32;                // + one producer warpgroup
  static constexpr int CTAS_PER_SM = BN == TILE ? 1 : 2;
  static constexpr int B_TILE_BYTES = BN * GEMM_BK * 8;
  static constexpr int STAGE_BYTES = GEMM_TILE_BYTES + B_TILE_BYTES;
};
// dynamic shared memory: ring + alignment slack + barriers + epilogue scratch (EPI_GRAD: xj [d][128] + red [d+2][2][BN];
// EPI_SUMSQ: red [2][BN])
__host__ __device__ constexpr int gemm_scratch_bytes(int epi, int bn, int d) {
  return epi == 2 ? (d * TILE + (d + 2) * 2 * bn) * 8 : (epi == 1 ? 2 * bn * 8 : 0);
}
__host__ __device__ constexpr int gemm_smem_bytes(int epi, int bn, int stages, int d) {
  return stages * (GEMM_TILE_BYTES + bn * GEMM_BK * 8) + 1024 /*align slack*/ + 256 /*barriers*/ + gemm_scratch_bytes(epi, bn, d);
}
constexpr int GEMM_SMEM_LIMIT_1 = 227 * 1024;          // one CTA per SM
constexpr int GEMM_SMEM_LIMIT_2 = 113 * 1024;          // two CTAs per SM: (228 KB - 2 x 1 KB reserved) / 2

enum KRange { KR_FULL = 0, KR_LE_M = 1, KR_GE_N = 2, KR_GE_MAX = 3, KR_GE_M = 4 };
enum TileSkip { TS_ALL = 0, TS_LOWER = 1 };
enum Raster { RO_M_FASTEST = 0, RO_HEAVY_FIRST = 1 };
enum Epilogue { EPI_STORE = 0, EPI_SUMSQ = 1, EPI_GRAD = 2 };

// A batch index z decomposes as z = node * zdiv + set: `set` selects the hyper-parameter set (3rd TMA coordinate, bat_z)
// and `node` an independent sub-problem inside one matrix (the same-level nodes of the recursive triangular inverse;
// row_z / col_z are its element strides).  Tile (0,0) of batch z starts at (col0 + node*col_z, row0 + node*row_z,
// bat0 + set*bat_z).
struct GemmOperand {
  int row0, col0, row_z, col_z, bat0, bat_z;
};

struct GemmParams {
  int tiles_m, tiles_n, kblocks;  // tiles_m, kblocks in units of 128; tiles_n in units of the kernel's BN (launch_gemm converts)
  int krange, tile_skip, raster;
  int batch;   // number of batches z (= nodes * zdiv)
  int zdiv;    // hyper-parameter sets per node (0: batch, i.e. a single node)
  int c_row_z, c_col_z, ct_row_z, ct_col_z;  // per-node element offsets of the EPI_STORE outputs
  int ksplit;  // >= 1: grid.y slices of the k range (EPI_GRAD only: its reductions are linear in the accumulator);
               // slice s of row block m writes partial row m*ksplit + s
  int serpentine;  // deal the work items to the CTAs in serpentine rounds (see gemm_item)
  int prefetch_kt; // EPI_GRAD: bulk-prefetch the tile's K* / H* entries into L2 at the start of the work item
  int tri_diag;    // the A operand is triangular inside its diagonal 128-block (k block == row block): 1 lower (zero for
                   // k > row: L^-1 under KR_LE_M), 2 upper (zero for k < row: L^-T under KR_GE_M).  The 8-row fragments that
                   // lie entirely in the zero part are skipped (they add exact zeros): the diagonal block costs ~56 % of a
                   // full one, which at n = 2048 is 6 % of a triangular product
  GemmOperand a, b;
  // EPI_STORE
  double alpha, beta;
  double* C;
  long long ldc, c_off_z;  // C element (r,c) of batch z at C[c_off_z*z + (c_row0+r)*ldc + c_col0 + c]
  int c_row0, c_col0;
  double* Ct;  // optional transposed copy: Ct[ct_off_z*z + (ct_row0+c)*ldct + ct_col0 + r]
  long long ldct, ct_off_z;
  int ct_row0, ct_col0;
  int sym_diag;  // diagonal tiles: C keeps c<=r, Ct keeps c<r (exactly symmetric result)
  // EPI_SUMSQ / EPI_GRAD: partial[((z*tiles_m + tile_m)*nvals + v)*ldp + tile_n*128 + col]
  double* partial;
  long long ldp;
  // EPI_SUMSQ, optional: the product itself, transposed, as the next product's K-major operand  Tt[z][col][row]
  double* Tt;
  long long tt_ld, tt_z;   // row pitch (= n_pad) and per-batch stride of Tt
  // EPI_GRAD: partial values per column are [S0 = sum_j b_j k_j, H0 = sum_j b_j h_j, P_q = sum_j b_j h_j xc_jq]
  int d;
  const double* Xtc;    // centred training inputs, SoA [d][ldx] (zero on the padding rows)
  long long ldx;
  const double* Kt;     // K*^T chunk [z][nc_ld][n_pad] written by K1 (also the B operand)
  const double* Ht;     // (dK/dr)/r chunk, same layout; null => RBF (h = -k)
  long long nc_ld, n_pad;
  // Two CTAs per SM (BN = 64): the second CTA to arrive on an SM walks ITS work items in reverse order.  The item lists are
  // sorted by weight and the serpentine deal gives co-resident CTAs near-identical weight sequences, so half of their
  // epilogues used to coincide and left the DMMA pipe idle; heavy-first against light-first they never meet.  Which CTA
  // computes an item, and therefore every result bit, is unchanged.  Null: every CTA walks forward.
  int* sm_arrivals;     // [#SMs] arrival counters (never reset: only the parity of consecutive arrivals matters)
};

// One unit of work of the persistent kernel: an output tile (tile_m, tile_n) of batch z, optionally one k slice of it.
// tile_n counts BN-wide column tiles.
struct GemmTile {
  int tile_m, tile_n, z, split, kb_begin, nk;
};

// Work items are numbered tile-fastest (row blocks sharing a B tile are neighbours), then k slice, then batch; with
// RO_HEAVY_FIRST the row blocks come in order of decreasing k range and, inside one row block, all batches and column
// tiles (so the whole list is sorted by weight).  With TS_LOWER only the tiles_m (tiles_m + 1) / 2 lower tiles are
// numbered (BN = 128 only).
__device__ __forceinline__ long long gemm_work_items(const GemmParams& p) {
  const long long tiles = p.tile_skip == TS_LOWER ? (long long)p.tiles_m * (p.tiles_m + 1) / 2 : (long long)p.tiles_m * p.tiles_n;
  return tiles * (p.ksplit > 1 ? p.ksplit : 1) * p.batch;
}

// The j-th work item of CTA b out of G: rounds of G items, every other round dealt in reverse (serpentine).  Where the
// item list is sorted by weight (triangular k ranges, heaviest first) this pairs each CTA's heavy items with light ones:
// at n = 2048 (row-block weights 16 ... 1, 74 column tiles, 148 CTAs) every CTA gets 68 k-blocks instead of 72 / 64.
__device__ __forceinline__ long long gemm_item(const GemmParams& p, long long j, int b, int G) {
  return j * G + ((p.serpentine && (j & 1)) ? G - 1 - b : b);
}

template <int BN>
__device__ __forceinline__ GemmTile gemm_decode(const GemmParams& p, long long w) {
  GemmTile t;
  const int ksplit = p.ksplit > 1 ? p.ksplit : 1;
  if (p.raster == RO_HEAVY_FIRST && p.tile_skip == TS_ALL) {
    // largest k range first: the last row block under KR_LE_M, the first under KR_GE_M
    const long long per_row = (long long)p.tiles_n * p.batch;
    const int rank = (int)(w / per_row), rem = (int)(w % per_row);
    t.z = rem / p.tiles_n;
    t.tile_n = rem % p.tiles_n;
    t.split = 0;
    t.tile_m = p.krange == KR_GE_M ? rank : p.tiles_m - 1 - rank;
  } else {
    const long long tiles = p.tile_skip == TS_LOWER ? (long long)p.tiles_m * (p.tiles_m + 1) / 2 : (long long)p.tiles_m * p.tiles_n;
    const int x = (int)(w % tiles);
    const int rest = (int)(w / tiles);
    t.split = rest % ksplit;
    t.z = rest / ksplit;
    if (p.tile_skip == TS_LOWER) {  // x -> (m, n), n <= m; rows with the longest k range first (KR_GE_MAX: small m)
      const int l = p.krange == KR_GE_MAX ? x : (int)tiles - 1 - x;
      int m = (int)((sqrt(8.0 * l + 1.0) - 1.0) * 0.5);
      while ((m + 1) * (m + 2) / 2 <= l) ++m;
      while (m * (m + 1) / 2 > l) --m;
      t.tile_m = m;
      t.tile_n = l - m * (m + 1) / 2;
    } else {
      t.tile_m = x % p.tiles_m;
      t.tile_n = x / p.tiles_m;
    }
  }
  const int nblk = (t.tile_n * BN) / TILE;  // the 128-block the column tile lies in
  int kb_begin = 0, kb_end = p.kblocks;
  if (p.krange == KR_LE_M) kb_end = min(t.tile_m + 1, p.kblocks);
  else if (p.krange == KR_GE_N) kb_begin = nblk;
  else if (p.krange == KR_GE_MAX) kb_begin = max(t.tile_m, nblk);
  else if (p.krange == KR_GE_M) kb_begin = t.tile_m;
  if (ksplit > 1) {
    const int len = (kb_end - kb_begin + ksplit - 1) / ksplit;
    kb_begin = min(kb_end, kb_begin + t.split * len);
    kb_end = min(kb_end, kb_begin + len);
  }
  t.kb_begin = kb_begin;
  t.nk = (kb_end - kb_begin) * (TILE / GEMM_BK);
  return t;
}

// Persistent: gridDim.x CTAs (CTAS_PER_SM per SM) walk the work items.  The producer thread runs ahead into the next
// item's first stages while the consumers are still in the previous item's epilogue, so CTA start-up, the first TMA
// round trip and the pipeline drain are paid once per CTA instead of once per tile.
template <int EPI, int BN, int STAGES>
__global__ void __launch_bounds__(GemmShape<BN>::THREADS, GemmShape<BN>::CTAS_PER_SM)
dgemm_nt_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const GemmParams p) {
  using Sh = GemmShape<BN>;
  constexpr int CW = Sh::CW, STAGE_BYTES = Sh::STAGE_BYTES;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // keep the pointer derived from the __shared__ symbol (so fragment loads compile to LDS, not generic LD)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full_bar = (uint64_t*)(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint8_t* scratch = smem + STAGES * STAGE_BYTES + 256;  // epilogue scratch, separate from the ring

  const long long n_work = gemm_work_items(p);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = (int)gridDim.x, cta = (int)blockIdx.x;
  int* rev_flag = reinterpret_cast<int*>(empty_bar + STAGES);   // (the 256-byte barrier block has room)
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], CW);
    }
    mbar_fence_init();
    int rev = 0;
    if (Sh::CTAS_PER_SM > 1 && p.sm_arrivals) {
      unsigned smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      rev = atomicAdd(p.sm_arrivals + smid, 1) & 1;
    }
    *rev_flag = rev;
  }
  __syncthreads();
  // this CTA's items are j = 0 .. n_mine - 1 of gemm_item; walked backwards when `reversed`
  const bool reversed = *rev_flag != 0;
  long long n_mine;
  {
    const long long rounds = n_work / G, rest = n_work % G;
    const long long off = (p.serpentine && (rounds & 1)) ? G - 1 - cta : cta;
    n_mine = rounds + (off < rest ? 1 : 0);
  }

  if (warp >= CW) {
    // ===================== TMA producer warpgroup (one elected thread issues) =====================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == CW && lane == 0) {
      tma_prefetch_desc(&mapA);
      tma_prefetch_desc(&mapB);
      long long g = 0;  // k iterations issued by this CTA so far (ring position)
      for (long long j = 0; j < n_mine; ++j) {
        const long long w = gemm_item(p, reversed ? n_mine - 1 - j : j, cta, G);
        const GemmTile t = gemm_decode<BN>(p, w);
        const int zdiv = p.zdiv > 0 ? p.zdiv : p.batch, node = t.z / zdiv, set = t.z % zdiv;
        const int a_row = p.a.row0 + node * p.a.row_z + t.tile_m * TILE, a_col = p.a.col0 + node * p.a.col_z;
        const int b_row = p.b.row0 + node * p.b.row_z + t.tile_n * BN, b_col = p.b.col0 + node * p.b.col_z;
        const int a_bat = p.a.bat0 + set * p.a.bat_z, b_bat = p.b.bat0 + set * p.b.bat_z;
        if constexpr (EPI == EPI_GRAD) {
          // The epilogue re-reads the K* (and H*) entries of this tile.  When K* is not this product's own B operand (the
          // second triangular product streams t = L^-1 K* instead) those lines have long left the L2: pull them back now,
          // a whole main loop ahead of their use (one 1 KB bulk prefetch per candidate row)
          if (p.prefetch_kt) {
            const long long off = ((long long)t.z * p.nc_ld + (long long)t.tile_n * BN) * p.n_pad + (long long)t.tile_m * TILE;
            for (int c = 0; c < BN; ++c) {
              l2_prefetch_bulk(p.Kt + off + (long long)c * p.n_pad, TILE * 8);
              if (p.Ht) l2_prefetch_bulk(p.Ht + off + (long long)c * p.n_pad, TILE * 8);
            }
          }
        }
        for (int it = 0; it < t.nk; ++it, ++g) {
          const int s = (int)(g % STAGES);
          if (g >= STAGES) mbar_wait(&empty_bar[s], (uint32_t)(((g / STAGES) & 1) ^ 1));
          mbar_arrive_expect_tx(&full_bar[s], STAGE_BYTES);
          const int k = t.kb_begin * TILE + it * GEMM_BK;
          uint8_t* st = smem + s * STAGE_BYTES;
          tma_load_3d(st, &mapA, &full_bar[s], a_col + k, a_row, a_bat);
          tma_load_3d(st + GEMM_TILE_BYTES, &mapB, &full_bar[s], b_col + k, b_row, b_bat);
        }
      }
    }
    return;
  }

  // ===================== consumers: DMMA main loop =====================
  if constexpr (BN == TILE) asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
  else asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
  // warp -> (wm, wn).  Warps w and w + 4 share an SM sub-partition (and its FP64 pipe): in the 8-warp shape one of them
  // takes the upper and one the lower 64 rows, so that the skipped fragments of a triangular diagonal block (tri_diag) are
  // spread evenly over the four pipes; in the 4-warp shape odd CTAs swap the halves (two CTAs share an SM).
  const int wm = BN == TILE ? (warp >> 2) * 64 : ((warp ^ cta) & 1) * 64;
  const int wn = BN == TILE ? (warp & 3) * 32 : (warp >> 1) * 32;
  const int g = lane >> 2, t = lane & 3;
  const int rr = frag_row(g);
  long long gi = 0;  // ring position, advances exactly like the producer's
  for (long long jw = 0; jw < n_mine; ++jw) {
  const long long w = gemm_item(p, reversed ? n_mine - 1 - jw : jw, cta, G);
  const GemmTile tl = gemm_decode<BN>(p, w);
  const int tile_m = tl.tile_m, tile_n = tl.tile_n, z = tl.z, split = tl.split, nk = tl.nk;
  const int ksplit = p.ksplit > 1 ? p.ksplit : 1;
  double acc[8][4][2];
#pragma unroll
  for (int f = 0; f < 8; ++f)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[f][j][0] = acc[f][j][1] = 0.0;

  for (int it = 0; it < nk; ++it, ++gi) {
    const int s = (int)(gi % STAGES);
    mbar_wait(&full_bar[s], (uint32_t)((gi / STAGES) & 1));
    const uint8_t* sA = smem + s * STAGE_BYTES + (wm + rr) * 128;
    const uint8_t* sB = smem + s * STAGE_BYTES + GEMM_TILE_BYTES + (wn + rr) * 128;
    if (p.tri_diag && tl.kb_begin + (it >> 3) == tile_m) {
      // diagonal block of a triangular A operand: fragment f (rows wm + 8 f .. + 7) of the k sub-step k0 .. k0 + 7 is all
      // zero when its rows lie above (lower) / below (upper) the k range
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int k0 = (it & 7) * GEMM_BK + 8 * h - wm;   // first k of the sub-step relative to the warp's first row
        const int f_lo = p.tri_diag == 1 ? max(0, k0 >> 3) : 0;
        const int f_hi = p.tri_diag == 2 ? min(8, k0 + 7 < 0 ? 0 : ((k0 + 7) >> 3) + 1) : 8;
        const int chunk = ((4 * h + t) ^ rr) << 4;
        double2 b[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = *reinterpret_cast<const double2*>(sB + j * 1024 + chunk);
#pragma unroll
        for (int f = 0; f < 8; ++f) {
          if (f >= f_lo && f < f_hi) {   // warp-uniform
            const double2 a = *reinterpret_cast<const double2*>(sA + f * 1024 + chunk);
#pragma unroll
            for (int j = 0; j < 4; ++j) dmma_884(acc[f][j][0], acc[f][j][1], a.x, b[j].x);
#pragma unroll
            for (int j = 0; j < 4; ++j) dmma_884(acc[f][j][0], acc[f][j][1], a.y, b[j].y);
          }
        }
      }
    } else {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int chunk = ((4 * h + t) ^ rr) << 4;  // SWIZZLE_128B: 16-byte chunk index XOR (row & 7)
      double2 a[8], b[4];
#pragma unroll
      for (int f = 0; f < 8; ++f) a[f] = *reinterpret_cast<const double2*>(sA + f * 1024 + chunk);
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = *reinterpret_cast<const double2*>(sB + j * 1024 + chunk);
#pragma unroll
      for (int f = 0; f < 8; ++f)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma_884(acc[f][j][0], acc[f][j][1], a[f].x, b[j].x);
#pragma unroll
      for (int f = 0; f < 8; ++f)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma_884(acc[f][j][0], acc[f][j][1], a[f].y, b[j].y);
    }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[s]);
  }

  // accumulator (f, j, e) of lane (g, t) is output element
  //   row = wm + 8 f + frag_row(g),   col = wn + 8 j + (e ? 4 + t : t)      (within the 128 x BN tile)
  if constexpr (EPI == EPI_STORE) {
    const bool diag = p.sym_diag && tile_m == (tile_n * BN) / TILE;
    const int zdiv = p.zdiv > 0 ? p.zdiv : p.batch, node = z / zdiv, set = z % zdiv;
    double* C = p.C ? p.C + p.c_off_z * set + (long long)node * p.c_row_z * p.ldc + node * p.c_col_z : nullptr;
    double* Ct = p.Ct ? p.Ct + p.ct_off_z * set + (long long)node * p.ct_row_z * p.ldct + node * p.ct_col_z : nullptr;
    // beta != 0: the tile of C is fetched in two batches of 32 independent loads per thread BEFORE any store of the
    // batch (interleaved load/store pairs would be kept in program order by the compiler -- the pointers may alias --
    // and cost one L2 round trip per element: 64 x ~0.3 us on the Cholesky's critical path)
#pragma unroll
    for (int fh = 0; fh < 2; ++fh) {
      double cin[4][4][2];
      if (C && p.beta != 0.0) {
#pragma unroll
        for (int f4 = 0; f4 < 4; ++f4) {
          const int r = tile_m * TILE + wm + 8 * (4 * fh + f4) + rr;
#pragma unroll
          for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int c = tile_n * BN + wn + 8 * j + (e ? 4 + t : t);
              cin[f4][j][e] = (!diag || c <= r) ? C[(long long)(p.c_row0 + r) * p.ldc + p.c_col0 + c] : 0.0;
            }
        }
      }
#pragma unroll
      for (int f4 = 0; f4 < 4; ++f4) {
        const int f = 4 * fh + f4;
        const int r = tile_m * TILE + wm + 8 * f + rr;
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int c = tile_n * BN + wn + 8 * j + (e ? 4 + t : t);
            double v = p.alpha * acc[f][j][e];
            if (C && (!diag || c <= r)) {
              if (p.beta != 0.0) v += p.beta * cin[f4][j][e];
              C[(long long)(p.c_row0 + r) * p.ldc + p.c_col0 + c] = v;
            }
            if (Ct && (!diag || c < r)) Ct[(long long)(p.ct_row0 + c) * p.ldct + p.ct_col0 + r] = v;
          }
      }
    }
  } else {
    // epilogue scratch lives past the ring: the producer may already be filling the ring for the next work item
    double* red = reinterpret_cast<double*>(scratch);  // EPI_SUMSQ: [2][BN]; EPI_GRAD: after the coordinates
    const int half = wm >> 6;
    if constexpr (EPI == EPI_SUMSQ) {
      if (p.Tt) {  // t = L^-1 k kept for the second triangular product (b = L^-T t): 64-byte runs per candidate row
        double* Tz = p.Tt + (long long)z * p.tt_z;
#pragma unroll
        for (int f = 0; f < 8; ++f) {
          const int r = tile_m * TILE + wm + 8 * f + rr;
#pragma unroll
          for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e)
              Tz[(long long)(tile_n * BN + wn + 8 * j + (e ? 4 + t : t)) * p.tt_ld + r] = acc[f][j][e];
        }
      }
      named_bar_sync(1, CW * 32);
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          double s2 = 0.0;
#pragma unroll
          for (int f = 0; f < 8; ++f) s2 = fma(acc[f][j][e], acc[f][j][e], s2);
          s2 += __shfl_xor_sync(0xffffffffu, s2, 4);
          s2 += __shfl_xor_sync(0xffffffffu, s2, 8);
          s2 += __shfl_xor_sync(0xffffffffu, s2, 16);
          if (g == 0) red[half * BN + wn + 8 * j + (e ? 4 + t : t)] = s2;
        }
      named_bar_sync(1, CW * 32);
      if (threadIdx.x < BN) {
        const int c = threadIdx.x;
        p.partial[((long long)(z * p.tiles_m + tile_m)) * p.ldp + tile_n * BN + c] = red[c] + red[BN + c];
      }
    } else {  // EPI_GRAD
      // b = W^-1 k sits in the accumulators.  The covariance k_ji and its radial derivative factor h_ji are NOT
      // recomputed: K1 left them in HBM/L2 (K*^T is this GEMM's own B operand), so each lane re-reads its 64
      // entries (8 consecutive j per candidate row: full 32-byte sectors) and reduces
      //   S0_i = sum_j b_ji k_ji        (posterior variance term k.b)
      //   H0_i = sum_j b_ji h_ji,  P_qi = sum_j b_ji h_ji xc_jq     (variance gradient, combined in K3)
      const int d = p.d;
      double* xj = reinterpret_cast<double*>(scratch);  // [d][128] centred training coordinates of this row block
      red = xj + d * TILE;                           // [(d+2)][2][BN]
      named_bar_sync(1, CW * 32);
      for (int idx = threadIdx.x; idx < d * TILE; idx += CW * 32)
        xj[idx] = p.Xtc[(idx / TILE) * p.ldx + tile_m * TILE + (idx % TILE)];
      const long long kt_off = ((long long)z * p.nc_ld + (long long)tile_n * BN) * p.n_pad + (long long)tile_m * TILE;
      const double* Kt = p.Kt + kt_off;
      const double* Ht = p.Ht ? p.Ht + kt_off : nullptr;
      double s0[4][2], h0[4][2];
#pragma unroll
      for (int j = 0; j < 4; ++j) s0[j][0] = s0[j][1] = h0[j][0] = h0[j][1] = 0.0;
      // the tile's K* (H*) entries are fetched in two batches of 32 (64) independent loads per lane, each issued in full
      // before its first use: one L2 / DRAM round trip per batch instead of one per fragment row (ncu: 58 % of this
      // epilogue's stall samples were long-scoreboard waits on eight dependent round trips)
#pragma unroll
      for (int fb = 0; fb < 2; ++fb) {
        double kv[4][4][2], hv[4][4][2];
#pragma unroll
        for (int f4 = 0; f4 < 4; ++f4) {
          const int r = wm + 8 * (4 * fb + f4) + rr;
#pragma unroll
          for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const long long o = (long long)(wn + 8 * j + (e ? 4 + t : t)) * p.n_pad + r;
              kv[f4][j][e] = __ldg(Kt + o);
              hv[f4][j][e] = Ht ? __ldg(Ht + o) : 0.0;
            }
        }
#pragma unroll
        for (int f4 = 0; f4 < 4; ++f4) {
          const int f = 4 * fb + f4;
#pragma unroll
          for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const double b = acc[f][j][e];
              s0[j][e] = fma(b, kv[f4][j][e], s0[j][e]);
              const double bh = Ht ? b * hv[f4][j][e] : -(b * kv[f4][j][e]);
              h0[j][e] += bh;
              acc[f][j][e] = bh;
            }
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          double v = s0[j][e], w2 = h0[j][e];
#pragma unroll
          for (int o = 4; o <= 16; o <<= 1) {
            v += __shfl_xor_sync(0xffffffffu, v, o);
            w2 += __shfl_xor_sync(0xffffffffu, w2, o);
          }
          if (g == 0) {
            red[(0 * 2 + half) * BN + wn + 8 * j + (e ? 4 + t : t)] = v;
            red[(1 * 2 + half) * BN + wn + 8 * j + (e ? 4 + t : t)] = w2;
          }
        }
      named_bar_sync(1, CW * 32);  // xj visible
      for (int q = 0; q < d; ++q) {
        double xjr[8];
#pragma unroll
        for (int f = 0; f < 8; ++f) xjr[f] = xj[q * TILE + wm + 8 * f + rr];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            double v = 0.0;
#pragma unroll
            for (int f = 0; f < 8; ++f) v = fma(acc[f][j][e], xjr[f], v);
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            if (g == 0) red[((q + 2) * 2 + half) * BN + wn + 8 * j + (e ? 4 + t : t)] = v;
          }
      }
      named_bar_sync(1, CW * 32);
      for (int idx = threadIdx.x; idx < (d + 2) * BN; idx += CW * 32) {
        const int v = idx / BN, c = idx % BN;
        p.partial[((long long)((z * p.tiles_m + tile_m) * ksplit + split) * (d + 2) + v) * p.ldp + tile_n * BN + c] =
            red[(v * 2) * BN + c] + red[(v * 2 + 1) * BN + c];
      }
    }
  }
  }  // work items
}

}  // namespace gpo
