// extern "C" surface of libocbo_b200 (declared in include/ocbo_b200.h) and the
// host-side blocked drivers (Cholesky, left triangular solve) that sequence the
// tile kernels.  No allocation, no synchronisation, no global state besides
// the per-kernel "max dynamic smem" opt-in done on first use.
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <mutex>
#include <unordered_map>

#include "common.cuh"
#include "kernels.h"

using namespace ocbo;

namespace ocbo {
unsigned long long g_launch_count = 0;
}
static thread_local char g_err[256] = "ok";

static int fail_arg(int k, const char* what) {
  snprintf(g_err, sizeof(g_err), "invalid argument %d: %s", k, what);
  return -k;
}
static int check(cudaError_t e, const char* where) {
  if (e == cudaSuccess) return 0;
  snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
  return 1000 + (int)e;
}
#define CK(call, where)              \
  do {                               \
    int rc_ = check((call), where);  \
    if (rc_) return rc_;             \
  } while (0)

static inline bool tile_ok(int n) { return n > 0 && n % TILE == 0; }

extern "C" {

unsigned long long ocbo_launch_count(void) {
  return __atomic_load_n(&ocbo::g_launch_count, __ATOMIC_RELAXED);
}
const char* ocbo_last_error_string(void) { return g_err; }
int ocbo_version(void) { return 100; }
int64_t ocbo_invdiag_elems(int n) { return (int64_t)(n / TILE) * TILE * TILE; }

int ocbo_gram(int kind, int dim, const double* X1, int n1, int64_t s_x1, const int32_t* nv1,
              const double* X2, int n2, int64_t s_x2, const int32_t* nv2, const double* theta,
              int64_t s_theta, double* K, int ldk, int64_t s_k, int batch, int flags,
              ocbo_stream_t stream) {
  if (kind < 0 || kind > 3) return fail_arg(1, "kernel kind");
  if (dim < 1 || dim > MAXD) return fail_arg(2, "dim must be in [1, 16]");
  if (!X1) return fail_arg(3, "X1 null");
  if (n1 <= 0 || n1 % 64) return fail_arg(4, "n1 must be a positive multiple of 64");
  if (!X2) return fail_arg(7, "X2 null");
  if (n2 <= 0 || n2 % 64) return fail_arg(8, "n2 must be a positive multiple of 64");
  if (!theta) return fail_arg(11, "theta null");
  if (!K) return fail_arg(13, "K null");
  if (ldk < n2 || (ldk & 1)) return fail_arg(14, "ldk must be even and >= n2");
  if (batch <= 0) return fail_arg(16, "batch");
  CK(launch_gram(kind, dim, X1, n1, s_x1, nv1, X2, n2, s_x2, nv2, theta, s_theta, K, ldk, s_k,
                 batch, flags, (cudaStream_t)stream),
     "ocbo_gram");
  return 0;
}

int ocbo_add_diag(double* A, int n, int lda, int64_t s_a, const int32_t* nv, const double* val,
                  int batch, ocbo_stream_t stream) {
  if (!A) return fail_arg(1, "A null");
  if (n <= 0) return fail_arg(2, "n");
  if (!val) return fail_arg(6, "val null");
  if (batch <= 0) return fail_arg(7, "batch");
  CK(launch_add_diag(A, n, lda, s_a, nv, val, batch, (cudaStream_t)stream), "ocbo_add_diag");
  return 0;
}

int ocbo_diag_max(const double* A, int n, int lda, int64_t s_a, const int32_t* nv, double* out,
                  int batch, ocbo_stream_t stream) {
  if (!A) return fail_arg(1, "A null");
  if (n <= 0) return fail_arg(2, "n");
  if (!out) return fail_arg(6, "out null");
  if (batch <= 0) return fail_arg(7, "batch");
  CK(launch_diag_max(A, n, lda, s_a, nv, out, batch, (cudaStream_t)stream), "ocbo_diag_max");
  return 0;
}

// Two-level blocked right-looking Cholesky.  Outer panels of `outer` columns
// (OCBO_POTRF_OUTER, default 2048) are factored with 128-wide steps; inside the
// panel the algorithm is LEFT-looking (block column j gets one rank-(j-J0)*128
// update before it is factored; OCBO_POTRF_INNER=right restores rank-128
// right-looking updates), and the trailing matrix is updated ONCE per outer
// panel with a rank-`outer` DMMA update, so all of the m^3/3 flops run with long
// k and every C tile is read and written once per (block column | outer panel).
// Measured at 16 trials in flight: left/2048 28.3 k samples/s, right/1024 27.4 k.
struct PotrfTimer {
  cudaEvent_t* ev; int* cls; int n; int cap;
  void mark(cudaStream_t st, int c) {
    if (!ev || n >= cap) return;
    cudaEventRecord(ev[n], st);
    cls[n] = c;
    ++n;
  }
};

// Look-ahead at the launch level (OCBO_POTRF_LOOKAHEAD; needs more than two outer panels): the rank-`outer` update of panel i is split into (a) the columns of the NEXT outer panel,
// (its first middle panel), on the caller's stream, and (b) everything to the right of it, on a
// helper stream.  The
// latency-bound chain of panel i+1 (diagonal blocks, panel solves, in-panel updates: a few CTAs
// at a time) then runs beside (b), which fills the SMs: one sweep trial alone 14.7 -> 14.1 ms
// (the 64 diagonal blocks and panel solves, 3.7 ms, stay a serial chain).  Per caller stream the library keeps ONE helper stream and two events, created
// on first use (the only state besides the per-kernel attributes).
struct Lookahead {
  cudaStream_t side = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
};
static Lookahead* lookahead_for(cudaStream_t st) {
  static std::mutex mu;
  static std::unordered_map<cudaStream_t, Lookahead*> table;
  std::lock_guard<std::mutex> lock(mu);
  auto it = table.find(st);
  if (it != table.end()) return it->second;
  Lookahead* la = new Lookahead();
  if (cudaStreamCreateWithFlags(&la->side, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&la->fork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&la->join, cudaEventDisableTiming) != cudaSuccess) {
    delete la;
    la = nullptr;
  }
  table[st] = la;
  return la;
}
// default: only for a single problem per launch (batch 1), where the chain's latency is the cost;
// with several trials per launch / in flight the other trials already fill the chain's gaps and
// the split update only adds launches (measured at 16 in flight: 30.1 k vs 30.5 k samples/s).
// OCBO_POTRF_LOOKAHEAD=1 forces it on, =0 off.
static bool lookahead_enabled(int batch) {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("OCBO_POTRF_LOOKAHEAD");
    v = e ? (atoi(e) != 0 ? 1 : 0) : 2;
  }
  return v == 2 ? batch == 1 : v == 1;
}

// outer panel width in tiles: OCBO_POTRF_OUTER, else 2048 columns; 1024 for factors below m = 6144 on the int8 path
static int potrf_outer_tiles(int nt, bool ozaki = false) {
  static int outer = -1;
  if (outer < 0) {
    const char* e = getenv("OCBO_POTRF_OUTER");
    int v = e ? atoi(e) : 0;
    outer = v < TILE ? 0 : v / TILE;
  }
  // (int8 path: 2048 again from m = 6144 on, now that the in-panel updates are int8 products too -- measured
  // 68.7 k against 66.8 k samples/s at m = 8192; smaller factors keep 1024 so that they have a trailing update at all)
  const int o = outer ? outer : ((ozaki && nt < 48) ? 1024 / TILE : 2048 / TILE);
  return o < nt ? o : nt;
}

// Trailing updates through the int8 tensor cores (ozaki.cu) when the caller hands a workspace to
// ocbo_potrf_ws: the finished outer panel is split into int8 slices once, then the rank-k update of the
// trailing matrix runs as tcgen05.mma kind::i8 products recombined in fp64 (6144 x 2048: 0.9 ms
// against 2.4 ms on the DMMA pipe).  Only worth it for long k and many tiles.
static bool ozaki_wanted(int kw, int R) {
  static int mink = -1;
  if (mink < 0) {
    const char* e = getenv("OCBO_OZAKI_MINK");
    mink = e ? atoi(e) : 512;
  }
  return kw >= mink && R >= 8;
}

// OCBO_OZAKI_GLOBAL (default 1): ONE split of the factor, with row scales fixed from the diagonal before the
// factorisation (|L_ik| <= sqrt(A_ii)), shared by the middle updates, the trailing updates and -- through
// ocbo_sample_tile_argmax_ws(l_in_ws) -- the sampling tail.  0: every update splits its own panel (row maxima).
static bool ozaki_global_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("OCBO_OZAKI_GLOBAL");
    v = e ? (atoi(e) != 0) : 1;
  }
  return v != 0;
}

// OCBO_OZAKI_INPANEL (needs the factor-wide image): shortest k of a left-looking in-panel update that runs on the
// int8 tensor cores too; the block columns are then split one by one as they are finished.  0 = never; default
// 128 = all of them (measured in one box: 65.4 k -> 67.4 k samples/s in flight, one trial alone 8.68 -> 8.41 ms).
static int ozaki_inpanel_mink() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("OCBO_OZAKI_INPANEL");
    v = e ? atoi(e) : 128;
  }
  return v;
}

static int potrf_run(double* A, int n, int lda, int64_t s_a, double* invdiag, int32_t* info,
                     int batch, cudaStream_t st, PotrfTimer* tm, const int32_t* nv = nullptr,
                     void* ws = nullptr, size_t ws_bytes = 0) {
  if (ws && ((lda & 3) || (s_a & 3) || ((uintptr_t)A & 31))) ws = nullptr;  // the split reads 256 bits at a time
  const int nt = n / TILE;
  const int64_t sInv = ocbo_invdiag_elems(n);
  const int NBO = potrf_outer_tiles(nt, ws != nullptr);
  static int left_env = -1;  // OCBO_POTRF_INNER=right|left forces the in-panel variant
  if (left_env < 0) {
    const char* e = getenv("OCBO_POTRF_INNER");
    left_env = !e ? 2 : (e[0] == 'r' ? 0 : 1);
  }
  // Left-looking panels run each block column's update as ONE long-k launch of
  // 2 (nt - j) batch CTAs: right for the sweep (the other streams fill the SMs), wrong for a
  // small problem on an idle GPU, where those few CTAs serialise (jointbran, m = 1280, one
  // problem: 19 launches x 27 us).  Small problems therefore take rank-128 right-looking
  // updates, whose launches expose (nt - j)^2 batch tiles of short k.
  const int left = left_env != 2 ? left_env : ((nt <= 16 && batch * nt * 2 < 296) ? 0 : 1);
  // Middle level (OCBO_POTRF_MID, in columns, default 1024; 0 disables): inside an outer
  // panel, sub-panels of MID columns are factored left-looking from their own first column, and
  // each finished sub-panel updates the REST of the outer panel's columns in one rank-MID launch
  // -- larger grids of medium k instead of long-k launches of few CTAs.  Measured (sweep shape):
  // one trial alone 17.0 -> 15.5 ms (14.8 with MID = 512), 2 streams x 2 trials +2 %, neutral
  // at the bench's 16 trials in flight, where the other streams fill the SMs anyway.
  // (with a workspace and 1024-wide outer panels the default is 512: the middle updates are int8 products too)
  static int mid_env = -2;
  if (mid_env == -2) {
    const char* e = getenv("OCBO_POTRF_MID");
    mid_env = e ? atoi(e) / TILE : -1;
  }
  const int mid = mid_env >= 0 ? mid_env : ((ws && NBO * TILE < 2048) ? 512 : 1024) / TILE;
  const int NBM = (mid > 0 && mid < NBO) ? mid : NBO;
  Lookahead* la = (!tm && nt > 2 * NBO && lookahead_enabled(batch)) ? lookahead_for(st) : nullptr;
  bool side_busy = false;  // a (b) update is in flight on the helper stream
  // factor-wide slice workspace (see ozaki_global_enabled)
  const bool glob = ws && !nv && nt > NBO && ozaki_global_enabled() && ozaki_wanted(NBO * TILE, nt - NBO) &&
                    ozaki_ws_bytes(n, n, batch) <= ws_bytes;
  if (glob) {
    CK(launch_ozaki_diag_scale(A, lda, s_a, n, ws, info, batch, st), "potrf ozaki row scales");
    if (tm) tm->mark(st, 4);
  }
  // in-panel updates on the int8 path: block columns are split as they are finished, not whole middle panels
  const int inp_mink = ozaki_inpanel_mink();
  const bool percol = glob && left && inp_mink > 0;
  // split the finished middle panel [m0, m1) (all its rows: the diagonal region too, the sampling tail reads it)
  auto split_mid = [&](int m0, int m1) -> int {
    if (percol) return 0;  // done column by column
    CK(launch_ozaki_split_fixed(A + (size_t)m0 * TILE * lda + (size_t)m0 * TILE, lda, s_a, (nt - m0) * TILE,
                                (m1 - m0) * TILE, m0 * TILE, m0 * (TILE / 32), n, ws, info, batch, st),
       "potrf ozaki split (factor-wide)");
    if (tm) tm->mark(st, 4);
    return 0;
  };
  for (int J0 = 0; J0 < nt; J0 += NBO) {
    const int J1 = (J0 + NBO < nt) ? J0 + NBO : nt;
    for (int j = J0; j < J1; ++j) {
      const int M0 = J0 + ((j - J0) / NBM) * NBM;  // first column of this middle panel
      if (left && j > M0) {
        // left-looking inside the (middle) panel: block column j receives the contributions of
        // all earlier columns of the panel in ONE rank-(j-M0)*128 update (long k, each C tile
        // read and written once) instead of j-M0 rank-128 updates
        NtParams u{};
        u.P = A + (size_t)j * TILE * lda + (size_t)M0 * TILE; u.ldp = lda; u.sP = s_a;
        u.Q = u.P; u.ldq = lda; u.sQ = s_a;
        u.C = A + (size_t)j * TILE * lda + (size_t)j * TILE; u.ldc = lda; u.sC = s_a;
        u.ntc = TILE / TN; u.ktiles = (j - M0) * (TILE / BK);
        u.tri = 0; u.lower_skip = 0; u.mode = 1; u.info = info;
        if (percol && (j - M0) * TILE >= inp_mink && nt - j >= 8) {
          CK(launch_ozaki_update_range(ws, n, M0 * (TILE / 32), (j - M0) * (TILE / 32), j * TILE, j * TILE, u.C, lda, s_a,
                                       nt - j, u.ntc, 0, 0, 0, 0, info, batch, st),
             "potrf panel update, ozaki");
          if (tm) tm->mark(st, 6);
        } else {
          CK(launch_gemm_nt(u, (nt - j) * u.ntc, batch, st), "potrf panel update");
          if (tm) tm->mark(st, 2);
        }
      }
      CK(launch_potrf_diag(A, lda, s_a, j * TILE, invdiag, sInv, info, batch, st, nv),
         "potrf_diag");
      if (tm) tm->mark(st, 0);
      const int r = nt - 1 - j;
      auto split_col = [&]() -> int {  // block column j, rows from its diagonal block down
        CK(launch_ozaki_split_fixed(A + (size_t)j * TILE * lda + (size_t)j * TILE, lda, s_a, (nt - j) * TILE, TILE,
                                    j * TILE, j * (TILE / 32), n, ws, info, batch, st),
           "potrf ozaki split (block column)");
        if (tm) tm->mark(st, 4);
        return 0;
      };
      if (r == 0) {
        if (percol) {
          int rc = split_col();
          if (rc) return rc;
        } else if (glob) {
          int rc = split_mid(M0, nt);
          if (rc) return rc;
        }
        break;
      }
      double* panel = A + (size_t)(j + 1) * TILE * lda + (size_t)j * TILE;
      NtParams p{};
      // panel solve: X <- X inv(L_jj)^T, in place (each CTA owns its 128 rows)
      p.P = panel; p.ldp = lda; p.sP = s_a;
      p.Q = invdiag + (size_t)j * TILE * TILE; p.ldq = TILE; p.sQ = sInv;
      p.C = panel; p.ldc = lda; p.sC = s_a;
      p.ntc = 1; p.ktiles = TILE / BK; p.mode = 0; p.info = info;
      CK(launch_gemm_nt(p, 2 * r, batch, st, 1), "potrf panel");  // 64-row tiles, full rows
      if (tm) tm->mark(st, 1);
      if (percol) {
        int rc = split_col();
        if (rc) return rc;
      }
      // inner update: the remaining columns of this outer panel, all rows below
      const int ncols = J1 - 1 - j;
      if (!left && ncols > 0) {
        NtParams u{};
        u.P = panel; u.ldp = lda; u.sP = s_a;
        u.Q = panel; u.ldq = lda; u.sQ = s_a;
        u.C = A + (size_t)(j + 1) * TILE * lda + (size_t)(j + 1) * TILE; u.ldc = lda; u.sC = s_a;
        u.ntc = ncols * (TILE / TN); u.ktiles = TILE / BK;
        u.row0 = (j + 1) * TILE; u.col0 = (j + 1) * TILE;
        u.tri = 0; u.lower_skip = 1; u.mode = 1; u.info = info;
        CK(launch_gemm_nt(u, r * u.ntc, batch, st), "potrf inner update");
        if (tm) tm->mark(st, 2);
      }
      // end of a middle panel [M0, M1): its columns update the rest of the outer panel
      const int M1 = (M0 + NBM < J1) ? M0 + NBM : J1;
      if (side_busy && j + 1 == M1) {
        // look-ahead: everything right of the first middle panel still belongs to update (b) of the
        // previous outer panel on the helper stream
        CK(cudaStreamWaitEvent(st, la->join, 0), "potrf lookahead join");
        side_busy = false;
      }
      if (glob && j + 1 == M1 && !(left && NBM < NBO && M1 < J1 && ozaki_wanted((M1 - M0) * TILE, nt - M1))) {
        int rc = split_mid(M0, M1);
        if (rc) return rc;
      }
      if (left && NBM < NBO && j + 1 == M1 && M1 < J1) {
        NtParams u{};
        u.P = A + (size_t)M1 * TILE * lda + (size_t)M0 * TILE; u.ldp = lda; u.sP = s_a;
        u.Q = u.P; u.ldq = lda; u.sQ = s_a;
        u.C = A + (size_t)M1 * TILE * lda + (size_t)M1 * TILE; u.ldc = lda; u.sC = s_a;
        u.ntc = (J1 - M1) * (TILE / TN); u.ktiles = (M1 - M0) * (TILE / BK);
        u.row0 = M1 * TILE; u.col0 = M1 * TILE;
        u.tri = 0; u.lower_skip = 1; u.mode = 1; u.info = info;
        const int kwm = (M1 - M0) * TILE, Rm = nt - M1;
        if (glob && ozaki_wanted(kwm, Rm)) {
          int rc = split_mid(M0, M1);
          if (rc) return rc;
          CK(launch_ozaki_update_range(ws, n, M0 * (TILE / 32), (M1 - M0) * (TILE / 32), M1 * TILE, M1 * TILE, u.C, lda, s_a,
                                       Rm, u.ntc, 0, 1, M1 * TILE, M1 * TILE, info, batch, st),
             "potrf middle update, ozaki (factor-wide)");
          if (tm) tm->mark(st, 6);
        } else if (!glob && ws && ozaki_wanted(kwm, Rm) && ozaki_ws_bytes(Rm * TILE, kwm, batch) <= ws_bytes) {
          CK(launch_ozaki_split(u.P, lda, s_a, Rm * TILE, kwm, ws, info, batch, st), "potrf middle ozaki split");
          if (tm) tm->mark(st, 4);
          CK(launch_ozaki_update(ws, Rm * TILE, kwm, 0, 0, u.C, lda, s_a, Rm, u.ntc, 0, 1, M1 * TILE, M1 * TILE, info,
                                 batch, st),
             "potrf middle update, ozaki");
          if (tm) tm->mark(st, 6);
        } else {
          CK(launch_gemm_nt(u, (nt - M1) * u.ntc, batch, st), "potrf middle update");
          if (tm) tm->mark(st, 2);
        }
      }
    }
    const int R = nt - J1;
    const int kw = (J1 - J0) * TILE;
    const bool ozg = glob && R > 0 && ozaki_wanted(kw, R);
    const bool oz = !glob && ws && R > 0 && ozaki_wanted(kw, R) && ozaki_ws_bytes(R * TILE, kw, batch) <= ws_bytes;
    const int kb0 = J0 * (TILE / 32), nkb = (J1 - J0) * (TILE / 32);
    if (oz) {
      // (the helper stream's update of the previous panel read the workspace: it was joined at the end of this
      // panel's first middle panel, or is joined here)
      if (side_busy) {
        CK(cudaStreamWaitEvent(st, la->join, 0), "potrf lookahead join");
        side_busy = false;
      }
      CK(launch_ozaki_split(A + (size_t)J1 * TILE * lda + (size_t)J0 * TILE, lda, s_a, R * TILE, kw, ws, info, batch, st),
         "potrf ozaki split");
      if (tm) tm->mark(st, 4);
    }
    if (R > 0 && la) {
      // look-ahead: (b) A[J2:, J2:] -= L[J2:, J0:J1] L[J2:, J0:J1]^T on the helper stream ...
      // (J2 = end of the next outer panel's FIRST middle panel: the narrower (a) is, the sooner the
      // next chain starts and the more of the update runs beside it)
      const int J2 = (J1 + NBM < nt) ? J1 + NBM : nt;
      const int Rb = nt - J2;
      CK(cudaEventRecord(la->fork, st), "potrf lookahead fork");  // panel [J0, J1) is final
      if (side_busy) CK(cudaStreamWaitEvent(st, la->join, 0), "potrf lookahead join");  // (b) of the panel before
      if (Rb > 0) {
        CK(cudaStreamWaitEvent(la->side, la->fork, 0), "potrf lookahead fork wait");
        NtParams u{};
        u.P = A + (size_t)J2 * TILE * lda + (size_t)J0 * TILE; u.ldp = lda; u.sP = s_a;
        u.Q = u.P; u.ldq = lda; u.sQ = s_a;
        u.C = A + (size_t)J2 * TILE * lda + (size_t)J2 * TILE; u.ldc = lda; u.sC = s_a;
        u.ntc = Rb * (TILE / TN); u.ktiles = (J1 - J0) * (TILE / BK); u.tri = 1; u.mode = 1;
        u.info = info;
        if (ozg)
          CK(launch_ozaki_update_range(ws, n, kb0, nkb, J2 * TILE, J2 * TILE, u.C, lda, s_a, Rb, 0, 1, 0, 0, 0, info, batch,
                                       la->side),
             "potrf outer update (b), ozaki (factor-wide)");
        else if (oz)
          CK(launch_ozaki_update(ws, R * TILE, kw, (J2 - J1) * TILE, (J2 - J1) * TILE, u.C, lda, s_a, Rb, 0, 1, 0, 0, 0,
                                 info, batch, la->side),
             "potrf outer update (b), ozaki");
        else
          CK(launch_gemm_nt(u, Rb * (Rb + 1), batch, la->side), "potrf outer update (b)");
        CK(cudaEventRecord(la->join, la->side), "potrf lookahead join record");
      }
      side_busy = Rb > 0;
      // ... and (a) the next outer panel's columns A[J1:, J1:J2] on the caller's stream
      NtParams u{};
      u.P = A + (size_t)J1 * TILE * lda + (size_t)J0 * TILE; u.ldp = lda; u.sP = s_a;
      u.Q = u.P; u.ldq = lda; u.sQ = s_a;
      u.C = A + (size_t)J1 * TILE * lda + (size_t)J1 * TILE; u.ldc = lda; u.sC = s_a;
      u.ntc = (J2 - J1) * (TILE / TN); u.ktiles = (J1 - J0) * (TILE / BK);
      u.row0 = J1 * TILE; u.col0 = J1 * TILE;
      u.tri = 0; u.lower_skip = 1; u.mode = 1; u.info = info;
      if (ozg)
        CK(launch_ozaki_update_range(ws, n, kb0, nkb, J1 * TILE, J1 * TILE, u.C, lda, s_a, R, u.ntc, 0, 1, J1 * TILE,
                                     J1 * TILE, info, batch, st),
           "potrf outer update (a), ozaki (factor-wide)");
      else if (oz)
        CK(launch_ozaki_update(ws, R * TILE, kw, 0, 0, u.C, lda, s_a, R, u.ntc, 0, 1, J1 * TILE, J1 * TILE, info, batch, st),
           "potrf outer update (a), ozaki");
      else
        CK(launch_gemm_nt(u, R * u.ntc, batch, st), "potrf outer update (a)");
    } else if (R > 0) {
      // outer update: A[J1:, J1:] -= L[J1:, J0:J1] L[J1:, J0:J1]^T  (lower tiles)
      NtParams u{};
      u.P = A + (size_t)J1 * TILE * lda + (size_t)J0 * TILE; u.ldp = lda; u.sP = s_a;
      u.Q = u.P; u.ldq = lda; u.sQ = s_a;
      u.C = A + (size_t)J1 * TILE * lda + (size_t)J1 * TILE; u.ldc = lda; u.sC = s_a;
      u.ntc = R * (TILE / TN); u.ktiles = (J1 - J0) * (TILE / BK); u.tri = 1; u.mode = 1;
      u.info = info;
      if (ozg)
        CK(launch_ozaki_update_range(ws, n, kb0, nkb, J1 * TILE, J1 * TILE, u.C, lda, s_a, R, 0, 1, 0, 0, 0, info, batch, st),
           "potrf outer update, ozaki (factor-wide)");
      else if (oz)
        CK(launch_ozaki_update(ws, R * TILE, kw, 0, 0, u.C, lda, s_a, R, 0, 1, 0, 0, 0, info, batch, st),
           "potrf outer update, ozaki");
      else
        CK(launch_gemm_nt(u, R * (R + 1), batch, st), "potrf outer update");  // 2:1 lower tiles
      if (tm) tm->mark(st, (oz || ozg) ? 5 : 3);
    }
  }
  if (side_busy) CK(cudaStreamWaitEvent(st, la->join, 0), "potrf lookahead final join");
  return 0;
}

int ocbo_potrf(double* A, int n, int lda, int64_t s_a, double* invdiag, int32_t* info, int batch,
               ocbo_stream_t stream) {
  if (!A) return fail_arg(1, "A null");
  if (!tile_ok(n)) return fail_arg(2, "n must be a positive multiple of 128");
  if (lda < n || (lda & 1)) return fail_arg(3, "lda must be even and >= n");
  if (!invdiag) return fail_arg(5, "invdiag null");
  if (!info) return fail_arg(6, "info null");
  if (batch <= 0) return fail_arg(7, "batch");
  return potrf_run(A, n, lda, s_a, invdiag, info, batch, (cudaStream_t)stream, nullptr);
}

int64_t ocbo_potrf_ws_bytes(int n, int batch) {
  if (!tile_ok(n) || batch <= 0) return 0;
  const int nt = n / TILE, NBO = potrf_outer_tiles(nt, true);
  if (nt <= NBO || !ozaki_wanted(NBO * TILE, nt - NBO)) return 0;
  if (ozaki_global_enabled()) return (int64_t)ozaki_ws_bytes(n, n, batch);  // the whole factor's slices
  return (int64_t)ozaki_ws_bytes((nt - NBO) * TILE, NBO * TILE, batch);
}

int ocbo_potrf_ws(double* A, int n, int lda, int64_t s_a, double* invdiag, int32_t* info, int batch,
                  void* ws, int64_t ws_bytes, ocbo_stream_t stream) {
  if (!A) return fail_arg(1, "A null");
  if (!tile_ok(n)) return fail_arg(2, "n must be a positive multiple of 128");
  if (lda < n || (lda & 1)) return fail_arg(3, "lda must be even and >= n");
  if (!invdiag) return fail_arg(5, "invdiag null");
  if (!info) return fail_arg(6, "info null");
  if (batch <= 0) return fail_arg(7, "batch");
  if (ws && (((uintptr_t)ws & 255) || ws_bytes < 0)) return fail_arg(8, "workspace must be 256-byte aligned");
  return potrf_run(A, n, lda, s_a, invdiag, info, batch, (cudaStream_t)stream, nullptr, nullptr, ws,
                   ws ? (size_t)ws_bytes : 0);
}

int ocbo_potrf_nv(double* A, int n, int lda, int64_t s_a, const int32_t* nv, double* invdiag,
                  int32_t* info, int batch, ocbo_stream_t stream) {
  if (!A) return fail_arg(1, "A null");
  if (!tile_ok(n)) return fail_arg(2, "n must be a positive multiple of 128");
  if (lda < n || (lda & 1)) return fail_arg(3, "lda must be even and >= n");
  if (!invdiag) return fail_arg(6, "invdiag null");
  if (!info) return fail_arg(7, "info null");
  if (batch <= 0) return fail_arg(8, "batch");
  return potrf_run(A, n, lda, s_a, invdiag, info, batch, (cudaStream_t)stream, nullptr, nv);
}

int ocbo_ladder_restore(double* A, const double* A0, int n, int lda, int64_t s_a, int lda0,
                        int64_t s_a0, const int32_t* nv, const int32_t* info, double factor,
                        int batch, ocbo_stream_t stream) {
  if (!A) return fail_arg(1, "A null");
  if (!A0) return fail_arg(2, "A0 null");
  if (!tile_ok(n)) return fail_arg(3, "n must be a positive multiple of 128");
  if (lda < n) return fail_arg(4, "lda");
  if (lda0 < n) return fail_arg(6, "lda0");
  if (!info) return fail_arg(9, "info null");
  if (!(factor >= 0.0)) return fail_arg(10, "factor must be >= 0");
  if (batch <= 0) return fail_arg(11, "batch");
  CK(launch_ladder_restore(A, A0, n, lda, s_a, lda0, s_a0, nv, info, factor, batch,
                           (cudaStream_t)stream),
     "ocbo_ladder_restore");
  return 0;
}

int ocbo_ladder_advance(int32_t* info, int32_t* rung, int p, int finish, int batch,
                        ocbo_stream_t stream) {
  if (!info) return fail_arg(1, "info null");
  if (!finish && !rung) return fail_arg(2, "rung null");
  if (batch <= 0) return fail_arg(5, "batch");
  CK(launch_ladder_advance(info, rung, p, finish, batch, (cudaStream_t)stream),
     "ocbo_ladder_advance");
  return 0;
}

int ocbo_ladder_adopt(double* A, const double* A1, int n, int lda, int64_t s_a, double* invdiag,
                      const double* invdiag1, int32_t* info, const int32_t* info1, int32_t* rung,
                      int p, int batch, ocbo_stream_t stream) {
  if (!A) return fail_arg(1, "A null");
  if (!A1) return fail_arg(2, "A1 null");
  if (!tile_ok(n)) return fail_arg(3, "n must be a positive multiple of 128");
  if (lda < n) return fail_arg(4, "lda");
  if (!invdiag || !invdiag1) return fail_arg(6, "invdiag null");
  if (!info || !info1) return fail_arg(8, "info null");
  if (!rung) return fail_arg(10, "rung null");
  if (batch <= 0) return fail_arg(12, "batch");
  CK(launch_ladder_adopt(A, A1, n, lda, s_a, invdiag, invdiag1, ocbo_invdiag_elems(n), info, info1,
                         rung, p, batch, (cudaStream_t)stream),
     "ocbo_ladder_adopt");
  return 0;
}

// Same factorisation, instrumented: CUDA events after every launch on `stream`,
// summed per kernel class into ms_host[4] = {diag, panel, inner update, outer
// update}.  Synchronous (bench.py's roofline leg only).
static int potrf_profile_run(double* A, int n, int lda, int64_t s_a, double* invdiag, int32_t* info, int batch,
                             void* ws, size_t ws_bytes, cudaStream_t st, float* ms_host, int classes) {
  const int cap = 6 * (n / TILE) + 2;
  cudaEvent_t* ev = new cudaEvent_t[cap];
  int* cls = new int[cap];
  for (int i = 0; i < cap; ++i) cudaEventCreate(&ev[i]);
  PotrfTimer tm{ev, cls, 0, cap};
  tm.mark(st, -1);
  int rc = potrf_run(A, n, lda, s_a, invdiag, info, batch, st, &tm, nullptr, ws, ws_bytes);
  cudaError_t se = cudaStreamSynchronize(st);
  for (int c = 0; c < classes; ++c) ms_host[c] = 0.f;
  if (!rc && se == cudaSuccess) {
    for (int i = 1; i < tm.n; ++i) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, ev[i - 1], ev[i]);
      if (cls[i] >= 0 && cls[i] < classes) ms_host[cls[i]] += ms;
    }
  }
  for (int i = 0; i < cap; ++i) cudaEventDestroy(ev[i]);
  delete[] ev;
  delete[] cls;
  if (rc) return rc;
  return check(se, "ocbo_potrf_profile sync");
}

int ocbo_potrf_profile(double* A, int n, int lda, int64_t s_a, double* invdiag, int32_t* info,
                       int batch, ocbo_stream_t stream, float* ms_host) {
  if (!A) return fail_arg(1, "A null");
  if (!tile_ok(n)) return fail_arg(2, "n must be a positive multiple of 128");
  if (!invdiag || !info || !ms_host) return fail_arg(5, "null");
  return potrf_profile_run(A, n, lda, s_a, invdiag, info, batch, nullptr, 0, (cudaStream_t)stream, ms_host, 4);
}

// ocbo_potrf_ws instrumented: ms_host[7] = {diag, panel, in-panel fp64 update, outer fp64 update, ozaki split,
// ozaki outer update, ozaki middle update}
int ocbo_potrf_ws_profile(double* A, int n, int lda, int64_t s_a, double* invdiag, int32_t* info, int batch,
                          void* ws, int64_t ws_bytes, ocbo_stream_t stream, float* ms_host) {
  if (!A) return fail_arg(1, "A null");
  if (!tile_ok(n)) return fail_arg(2, "n must be a positive multiple of 128");
  if (!invdiag || !info || !ms_host) return fail_arg(5, "null");
  if (ws && ((uintptr_t)ws & 255)) return fail_arg(8, "workspace must be 256-byte aligned");
  return potrf_profile_run(A, n, lda, s_a, invdiag, info, batch, ws, ws ? (size_t)ws_bytes : 0, (cudaStream_t)stream,
                           ms_host, 7);
}


int ocbo_trsm_lower(const double* L, int n, int ldl, int64_t s_l, const double* invdiag, double* B,
                    int nrhs, int ldb, int64_t s_b, const int32_t* info, int batch,
                    ocbo_stream_t stream) {
  if (!L) return fail_arg(1, "L null");
  if (!tile_ok(n)) return fail_arg(2, "n must be a positive multiple of 128");
  if (ldl < n || (ldl & 1)) return fail_arg(3, "ldl");
  if (!invdiag) return fail_arg(5, "invdiag null");
  if (!B) return fail_arg(6, "B null");
  if (!tile_ok(nrhs)) return fail_arg(7, "nrhs must be a positive multiple of 128");
  if (ldb < nrhs || (ldb & 1)) return fail_arg(8, "ldb");
  if (batch <= 0) return fail_arg(11, "batch");
  cudaStream_t st = (cudaStream_t)stream;
  const int nt = n / TILE, ntc = nrhs / TN;
  const int64_t sInv = ocbo_invdiag_elems(n);
  // Block forward substitution with the inverted diagonal blocks.  Left-looking (block row I
  // gets ONE update of k = 128 I, then the multiply by inv(L_II)) launches ntc * batch CTAs per
  // step: right for wide right-hand sides (the sweep: 128 column tiles per trial).  When that
  // does not fill the GPU (conth42: 50 contexts x 2 column tiles on 148 SMs, k up to 640 per
  // CTA) the right-looking form is used instead: finished block row I updates ALL rows below
  // it in one launch of (nt - 1 - I) ntc batch CTAs with k = 128.
  const bool right = nt >= 3 && (long long)ntc * batch < 296;
  for (int I = 0; I < nt; ++I) {
    double* BI = B + (size_t)I * TILE * ldb;
    if (!right && I > 0) {
      NnParams u{};
      u.P = L + (size_t)I * TILE * ldl; u.ldp = ldl; u.sP = s_l;
      u.R = B; u.ldr = ldb; u.sR = s_b;
      u.C = BI; u.ldc = ldb; u.sC = s_b;
      u.ktiles = I * (TILE / BK); u.ntr = 1; u.tri_k = 0; u.pair_rows = 0; u.mode = 1; u.info = info;
      CK(launch_gemm_nn(u, 1, ntc, batch, st), "trsm update");
    }
    NnParams m{};
    m.P = invdiag + (size_t)I * TILE * TILE; m.ldp = TILE; m.sP = sInv;
    m.R = BI; m.ldr = ldb; m.sR = s_b;
    m.C = BI; m.ldc = ldb; m.sC = s_b;
    m.ktiles = TILE / BK; m.ntr = 1; m.tri_k = 0; m.pair_rows = 0; m.mode = 0; m.info = info;
    CK(launch_gemm_nn(m, 1, ntc, batch, st), "trsm diag");
    if (right && I + 1 < nt) {
      NnParams u{};
      u.P = L + (size_t)(I + 1) * TILE * ldl + (size_t)I * TILE; u.ldp = ldl; u.sP = s_l;
      u.R = BI; u.ldr = ldb; u.sR = s_b;
      u.C = B + (size_t)(I + 1) * TILE * ldb; u.ldc = ldb; u.sC = s_b;
      u.ktiles = TILE / BK; u.ntr = 1; u.tri_k = 0; u.pair_rows = 0; u.mode = 1; u.info = info;
      CK(launch_gemm_nn(u, nt - 1 - I, ntc, batch, st), "trsm trailing update");
    }
  }
  return 0;
}

int ocbo_solve_alpha_lml(const double* L, int n, int ldl, int64_t s_l, const double* invdiag,
                         const double* y, int64_t s_y, const int32_t* nv, const double* theta,
                         int64_t s_theta, double* alpha, int64_t s_alpha, double* lml,
                         const int32_t* info, int batch, ocbo_stream_t stream) {
  if (!L) return fail_arg(1, "L null");
  if (!tile_ok(n)) return fail_arg(2, "n must be a positive multiple of 128");
  if (n > 24576) return fail_arg(2, "n too large for the single-CTA solve");
  if (!invdiag) return fail_arg(5, "invdiag null");
  if (!y) return fail_arg(6, "y null");
  if (!theta) return fail_arg(9, "theta null");
  if (batch <= 0) return fail_arg(15, "batch");
  CK(launch_solve_alpha_lml(L, n, ldl, s_l, invdiag, ocbo_invdiag_elems(n), y, s_y, nv, theta,
                            s_theta, alpha, s_alpha, lml, info, batch, (cudaStream_t)stream),
     "ocbo_solve_alpha_lml");
  return 0;
}

int ocbo_solve_alpha_w_lml(const double* L, int n, int ldl, int64_t s_l, const double* invdiag,
                           const double* y, int64_t s_y, const int32_t* nv, const double* theta,
                           int64_t s_theta, double* alpha, int64_t s_alpha, double* w, int64_t s_w,
                           double* lml, const int32_t* info, int batch, ocbo_stream_t stream) {
  if (!L) return fail_arg(1, "L null");
  if (!tile_ok(n)) return fail_arg(2, "n must be a positive multiple of 128");
  if (n > 24576) return fail_arg(2, "n too large for the single-CTA solve");
  if (!invdiag) return fail_arg(5, "invdiag null");
  if (!y) return fail_arg(6, "y null");
  if (!theta) return fail_arg(9, "theta null");
  if (!alpha && !w && !lml) return fail_arg(11, "no output requested");
  if (batch <= 0) return fail_arg(17, "batch");
  CK(launch_solve_alpha_lml(L, n, ldl, s_l, invdiag, ocbo_invdiag_elems(n), y, s_y, nv, theta,
                            s_theta, alpha, s_alpha, lml, info, batch, (cudaStream_t)stream, w, s_w),
     "ocbo_solve_alpha_w_lml");
  return 0;
}

int ocbo_post_mean(int kind, int dim, const double* Xs, int m, int64_t s_xs, const int32_t* mv,
                   const double* X, int n, int64_t s_x, const int32_t* nv, const double* theta,
                   int64_t s_theta, const double* alpha, int64_t s_alpha, double* mean,
                   int64_t s_mean, int batch, ocbo_stream_t stream) {
  if (kind < 0 || kind > 3) return fail_arg(1, "kernel kind");
  if (dim < 1 || dim > MAXD) return fail_arg(2, "dim");
  if (!Xs) return fail_arg(3, "Xs null");
  if (m <= 0) return fail_arg(4, "m");
  if (n < 0) return fail_arg(8, "n");
  if (n > 0 && (!X || !alpha)) return fail_arg(7, "X/alpha null");
  if (!theta) return fail_arg(11, "theta null");
  if (!mean) return fail_arg(15, "mean null");
  if (batch <= 0) return fail_arg(17, "batch");
  CK(launch_post_mean(kind, dim, Xs, m, s_xs, mv, X, n, s_x, nv, theta, s_theta, alpha, s_alpha,
                      mean, s_mean, batch, (cudaStream_t)stream),
     "ocbo_post_mean");
  return 0;
}

int ocbo_post_cov(int kind, int dim, const double* Xs, int m, int64_t s_xs, const int32_t* mv,
                  const double* V, int n, int ldv, int64_t s_v, const double* theta,
                  int64_t s_theta, double* Sigma, int lds, int64_t s_sigma, int lower_only,
                  int batch, ocbo_stream_t stream) {
  if (kind < 0 || kind > 3) return fail_arg(1, "kernel kind");
  if (dim < 1 || dim > MAXD) return fail_arg(2, "dim");
  if (!Xs) return fail_arg(3, "Xs null");
  if (!tile_ok(m)) return fail_arg(4, "m must be a positive multiple of 128");
  if (n < 0 || n % TILE) return fail_arg(8, "n must be a multiple of 128");
  if (n > 0 && !V) return fail_arg(7, "V null");
  if (n > 0 && (ldv < m || (ldv & 1))) return fail_arg(9, "ldv");
  if (!theta) return fail_arg(11, "theta null");
  if (!Sigma) return fail_arg(13, "Sigma null");
  if (lds < m || (lds & 1)) return fail_arg(14, "lds");
  if (batch <= 0) return fail_arg(17, "batch");
  const int nt = m / TILE;
  CovParams p{};
  p.Xa = Xs; p.sXa = s_xs; p.ma = m; p.mav = mv;
  p.Xb = Xs; p.sXb = s_xs; p.mb = m; p.mbv = mv;
  p.V1 = V; p.ldv1 = ldv; p.sV1 = s_v;
  p.V2 = V; p.ldv2 = ldv; p.sV2 = s_v;
  p.theta = theta; p.sTheta = s_theta;
  p.C = Sigma; p.ldc = lds; p.sC = s_sigma;
  p.ntc = m / TN; p.ktiles = n / BK; p.dim = dim; p.kind = kind; p.sym = 1; p.tri = lower_only ? 1 : 0;
  CK(launch_postcov(p, lower_only ? nt * (nt + 1) : nt * (m / TN), batch, (cudaStream_t)stream),
     "ocbo_post_cov");
  return 0;
}

// Sigma = K(Xs, Xs) - V^T V with the product on the int8 tensor cores: the Gram matrix is written first (lower
// tiles), V is split column-wise into int8 slices, and the fused update kernel subtracts V^T V from it.
static bool post_cov_ozaki_wanted(int m, int n) { return n >= 512 && m / TILE >= 8; }

int64_t ocbo_post_cov_ws_bytes(int m, int n, int batch) {
  if (!tile_ok(m) || n <= 0 || n % TILE || batch <= 0 || !post_cov_ozaki_wanted(m, n)) return 0;
  return (int64_t)ozaki_ws_bytes(m, n, batch);
}

int ocbo_post_cov_ws(int kind, int dim, const double* Xs, int m, int64_t s_xs, const int32_t* mv,
                     const double* V, int n, int ldv, int64_t s_v, const double* theta,
                     int64_t s_theta, double* Sigma, int lds, int64_t s_sigma, int lower_only,
                     int batch, void* ws, int64_t ws_bytes, ocbo_stream_t stream) {
  if (!ws || !lower_only || !tile_ok(m) || n <= 0 || n % TILE || !post_cov_ozaki_wanted(m, n) || batch <= 0 ||
      ws_bytes < (int64_t)ozaki_ws_bytes(m, n, batch))
    return ocbo_post_cov(kind, dim, Xs, m, s_xs, mv, V, n, ldv, s_v, theta, s_theta, Sigma, lds, s_sigma,
                         lower_only, batch, stream);
  if ((uintptr_t)ws & 255) return fail_arg(18, "workspace must be 256-byte aligned");
  if (!V) return fail_arg(7, "V null");
  if (ldv < m || (ldv & 1)) return fail_arg(9, "ldv");
  static int fuse_env = -1;  // OCBO_OZAKI_FUSE_GRAM=0: Gram written by ocbo_gram first (A/B switch)
  if (fuse_env < 0) {
    const char* e = getenv("OCBO_OZAKI_FUSE_GRAM");
    fuse_env = e ? (atoi(e) != 0) : 1;
  }
  const bool fuse = fuse_env && !mv && dim <= ozaki_gram_max_dim();  // Gram term in the update kernel's epilogue
  if (!fuse) {
    int rc = ocbo_gram(kind, dim, Xs, m, s_xs, mv, Xs, m, s_xs, mv, theta, s_theta, Sigma, lds, s_sigma, batch,
                       OCBO_GRAM_SYM | OCBO_GRAM_LOWER, stream);
    if (rc) return rc;
  }
  CK(launch_ozaki_split(V, ldv, s_v, m, n, ws, nullptr, batch, (cudaStream_t)stream, 1), "post_cov ozaki split");
  OzGram g{Xs, s_xs, dim, kind, theta, s_theta};
  CK(launch_ozaki_update(ws, m, n, 0, 0, Sigma, lds, s_sigma, m / TILE, 0, 1, 0, 0, 0, nullptr, batch,
                         (cudaStream_t)stream, fuse ? &g : nullptr),
     "post_cov ozaki update");
  return 0;
}

int ocbo_post_cross_cov(int kind, int dim, const double* Xa, int ma, int64_t s_xa,
                        const int32_t* mav, const double* Xb, int mb, int64_t s_xb,
                        const int32_t* mbv, const double* V1, int ldv1, int64_t s_v1,
                        const double* V2, int ldv2, int64_t s_v2, int n, const double* theta,
                        int64_t s_theta, double* C, int ldc, int64_t s_c, int batch,
                        ocbo_stream_t stream) {
  if (kind < 0 || kind > 3) return fail_arg(1, "kernel kind");
  if (dim < 1 || dim > MAXD) return fail_arg(2, "dim");
  if (!Xa || !Xb) return fail_arg(3, "Xa/Xb null");
  if (!tile_ok(ma)) return fail_arg(4, "ma must be a positive multiple of 128");
  if (!tile_ok(mb)) return fail_arg(8, "mb must be a positive multiple of 128");
  if (n < 0 || n % TILE) return fail_arg(17, "n must be a multiple of 128");
  if (n > 0 && (!V1 || !V2)) return fail_arg(11, "V1/V2 null");
  if (!theta) return fail_arg(18, "theta null");
  if (!C) return fail_arg(20, "C null");
  if (ldc < mb || (ldc & 1)) return fail_arg(21, "ldc");
  if (batch <= 0) return fail_arg(23, "batch");
  CovParams p{};
  p.Xa = Xa; p.sXa = s_xa; p.ma = ma; p.mav = mav;
  p.Xb = Xb; p.sXb = s_xb; p.mb = mb; p.mbv = mbv;
  p.V1 = V1; p.ldv1 = ldv1; p.sV1 = s_v1;
  p.V2 = V2; p.ldv2 = ldv2; p.sV2 = s_v2;
  p.theta = theta; p.sTheta = s_theta;
  p.C = C; p.ldc = ldc; p.sC = s_c;
  p.ntc = mb / TN; p.ktiles = n / BK; p.dim = dim; p.kind = kind; p.sym = 0; p.tri = 0;
  CK(launch_postcov(p, (ma / TILE) * (mb / TN), batch, (cudaStream_t)stream),
     "ocbo_post_cross_cov");
  return 0;
}

int ocbo_solve_forward_lml(const double* L, int n, int ldl, int64_t s_l, const double* invdiag,
                           const double* y, int64_t s_y, const int32_t* nv, const double* theta,
                           int64_t s_theta, double* w, int64_t s_w, double* lml,
                           const int32_t* info, int batch, ocbo_stream_t stream) {
  if (!L) return fail_arg(1, "L null");
  if (!tile_ok(n)) return fail_arg(2, "n must be a positive multiple of 128");
  if (n > 24576) return fail_arg(2, "n too large for the single-CTA solve");
  if (!invdiag) return fail_arg(5, "invdiag null");
  if (!y) return fail_arg(6, "y null");
  if (!theta) return fail_arg(9, "theta null");
  if (!w && !lml) return fail_arg(11, "w and lml both null");
  if (batch <= 0) return fail_arg(15, "batch");
  CK(launch_solve_alpha_lml(L, n, ldl, s_l, invdiag, ocbo_invdiag_elems(n), y, s_y, nv, theta,
                            s_theta, nullptr, 0, lml, info, batch, (cudaStream_t)stream, w, s_w),
     "ocbo_solve_forward_lml");
  return 0;
}

int ocbo_post_mean_var_from_v(const double* V, int n, int ldv, int64_t s_v, const int32_t* nv,
                              const double* w, int64_t s_w, const double* theta, int64_t s_theta,
                              int m, const int32_t* mv, double* mean, int64_t s_mean, double* var,
                              int64_t s_var, int batch, ocbo_stream_t stream) {
  if (n < 0) return fail_arg(2, "n");
  if (n > 0 && (!V || !w)) return fail_arg(1, "V / w null");
  if (!theta) return fail_arg(8, "theta null");
  if (m <= 0) return fail_arg(10, "m");
  if (!mean && !var) return fail_arg(12, "mean and var both null");
  if (batch <= 0) return fail_arg(16, "batch");
  CK(launch_post_var_ei(V, n, ldv, s_v, nv, nullptr, 0, theta, s_theta, nullptr, m, mv, var, s_var,
                        nullptr, 0, batch, (cudaStream_t)stream, w, s_w, mean, s_mean),
     "ocbo_post_mean_var_from_v");
  return 0;
}

int ocbo_post_var_ei(const double* V, int n, int ldv, int64_t s_v, const int32_t* nv,
                     const double* mean, int64_t s_mean, const double* theta, int64_t s_theta,
                     const double* best, int m, const int32_t* mv, double* var, int64_t s_var,
                     double* ei, int64_t s_ei, int batch, ocbo_stream_t stream) {
  if (n < 0) return fail_arg(2, "n");
  if (n > 0 && !V) return fail_arg(1, "V null");
  if (!theta) return fail_arg(8, "theta null");
  if (m <= 0) return fail_arg(11, "m");
  if (!var && !ei) return fail_arg(13, "var and ei both null");
  if (ei && (!best || !mean)) return fail_arg(10, "ei needs mean and best");
  if (batch <= 0) return fail_arg(17, "batch");
  CK(launch_post_var_ei(V, n, ldv, s_v, nv, mean, s_mean, theta, s_theta, best, m, mv, var, s_var,
                        ei, s_ei, batch, (cudaStream_t)stream),
     "ocbo_post_var_ei");
  return 0;
}

int ocbo_symmetrize_lower(double* A, int n, int lda, int64_t s_a, int batch,
                          ocbo_stream_t stream) {
  if (!A) return fail_arg(1, "A null");
  if (n <= 0 || n % 32) return fail_arg(2, "n must be a positive multiple of 32");
  if (batch <= 0) return fail_arg(5, "batch");
  CK(launch_symmetrize(A, n, lda, s_a, batch, (cudaStream_t)stream), "ocbo_symmetrize_lower");
  return 0;
}

int ocbo_philox_normal(double* Z, int m, int S, int ldz, int64_t s_z, uint64_t seed,
                       const int32_t* trial_ids, int batch, ocbo_stream_t stream) {
  if (!Z) return fail_arg(1, "Z null");
  if (m <= 0) return fail_arg(2, "m");
  if (S <= 0) return fail_arg(3, "S");
  if (ldz < S) return fail_arg(4, "ldz");
  if (batch <= 0) return fail_arg(8, "batch");
  CK(launch_philox_normal(Z, m, S, ldz, s_z, seed, trial_ids, batch, (cudaStream_t)stream),
     "ocbo_philox_normal");
  return 0;
}

int ocbo_sample(const double* mean, int64_t s_mean, const double* L, int m, int ldl, int64_t s_l,
                const double* Z, int S, int ldz, int64_t s_z, double* F, int ldf, int64_t s_f,
                const int32_t* info, int batch, ocbo_stream_t stream) {
  if (!mean) return fail_arg(1, "mean null");
  if (!L) return fail_arg(3, "L null");
  if (m <= 0) return fail_arg(4, "m");
  if (!Z) return fail_arg(7, "Z null");
  if (S <= 0) return fail_arg(8, "S");
  if (!F) return fail_arg(11, "F null");
  if (batch <= 0) return fail_arg(15, "batch");
  cudaStream_t st = (cudaStream_t)stream;
  if (S <= 8) {
    CK(launch_sample_rowdot(mean, s_mean, L, m, ldl, s_l, Z, S, ldz, s_z, F, ldf, s_f, info, batch,
                            st),
       "ocbo_sample rowdot");
    return 0;
  }
  if (!tile_ok(m)) return fail_arg(4, "m must be a multiple of 128 on the DMMA path");
  if (S % TN) return fail_arg(8, "S must be <= 8 or a multiple of 64");
  if ((ldz & 1) || (ldf & 1) || (ldl & 1)) return fail_arg(9, "leading dimensions must be even");
  NnParams p{};
  p.P = L; p.ldp = ldl; p.sP = s_l;
  p.R = Z; p.ldr = ldz; p.sR = s_z;
  p.C = F; p.ldc = ldf; p.sC = s_f;
  p.rowvec = mean; p.sRowvec = s_mean;
  p.ktiles = m / BK; p.ntr = m / TILE; p.tri_k = 1; p.pair_rows = 1; p.mode = 2; p.info = info;
  CK(launch_gemm_nn(p, (p.ntr + 1) / 2, S / TN, batch, st), "ocbo_sample dmma");
  return 0;
}

int ocbo_sample_tile_argmax(const double* mean, int64_t s_mean, const double* L, int m, int ldl,
                            int64_t s_l, const double* Z, int S, int ldz, int64_t s_z,
                            double* tile_max, int32_t* tile_arg, const int32_t* info, int batch,
                            ocbo_stream_t stream) {
  if (!mean) return fail_arg(1, "mean null");
  if (!L) return fail_arg(3, "L null");
  if (!tile_ok(m)) return fail_arg(4, "m must be a positive multiple of 128");
  if (ldl < m || (ldl & 1)) return fail_arg(5, "ldl must be even and >= m");
  if (!Z) return fail_arg(7, "Z null");
  if (S <= 0 || S % 32) return fail_arg(8, "S must be a positive multiple of 32");
  if (ldz < S || (ldz & 1)) return fail_arg(9, "ldz must be even and >= S");
  if (!tile_max || !tile_arg) return fail_arg(11, "outputs null");
  if (batch <= 0) return fail_arg(14, "batch");
  NnParams p{};
  p.P = L; p.ldp = ldl; p.sP = s_l;
  p.R = Z; p.ldr = ldz; p.sR = s_z;
  p.rowvec = mean; p.sRowvec = s_mean;
  p.ktiles = m / BK; p.ntr = m / TILE; p.tri_k = 1; p.pair_rows = 1; p.mode = 2; p.info = info;
  p.tile_max = tile_max; p.tile_arg = tile_arg; p.out_ld = S;
  CK(launch_sample_argmax(p, S / 32, batch, (cudaStream_t)stream), "ocbo_sample_tile_argmax");
  return 0;
}

// The same tail with the product on the int8 tensor cores (csrc/ozaki.cu): L and Z^T are split into int8 slices in
// the caller's workspace, the per-tile max / arg-max come out of the tensor-memory epilogue.
static bool sample_ozaki_wanted(int m, int S) { return m / TILE >= 8 && S % 64 == 0 && m <= 16384; }

int64_t ocbo_sample_ws_bytes(int m, int S, int batch) {
  if (!tile_ok(m) || S <= 0 || batch <= 0 || !sample_ozaki_wanted(m, S)) return 0;
  return (int64_t)ozaki_sample_ws_bytes(m, S, batch);
}

int ocbo_sample_tile_argmax_ws(const double* mean, int64_t s_mean, const double* L, int m, int ldl,
                               int64_t s_l, const double* Z, int S, int ldz, int64_t s_z,
                               double* tile_max, int32_t* tile_arg, const int32_t* info, int batch,
                               void* ws, int64_t ws_bytes, int l_in_ws, ocbo_stream_t stream) {
  if (!ws || !tile_ok(m) || S <= 0 || batch <= 0 || !sample_ozaki_wanted(m, S) || (ldl & 3) || (s_l & 3) ||
      ((uintptr_t)L & 31) || ws_bytes < (int64_t)ozaki_sample_ws_bytes(m, S, batch))
    return ocbo_sample_tile_argmax(mean, s_mean, L, m, ldl, s_l, Z, S, ldz, s_z, tile_max, tile_arg, info, batch,
                                   stream);
  if ((uintptr_t)ws & 255) return fail_arg(15, "workspace must be 256-byte aligned");
  if (!mean) return fail_arg(1, "mean null");
  if (!L) return fail_arg(3, "L null");
  if (ldl < m) return fail_arg(5, "ldl must be >= m");
  if (!Z) return fail_arg(7, "Z null");
  if (ldz < S || (ldz & 1)) return fail_arg(9, "ldz must be even and >= S");
  if (!tile_max || !tile_arg) return fail_arg(11, "outputs null");
  // l_in_ws: only meaningful when ocbo_potrf_ws would have left the factor's slices there (same switches)
  const bool reuse = l_in_ws && ozaki_global_enabled() && ocbo_potrf_ws_bytes(m, batch) == (int64_t)ozaki_ws_bytes(m, m, batch);
  CK(launch_ozaki_sample(mean, s_mean, L, m, ldl, s_l, Z, S, ldz, s_z, tile_max, tile_arg, info, batch, ws,
                         (cudaStream_t)stream, reuse ? 1 : 0),
     "ocbo_sample_tile_argmax_ws");
  return 0;
}

int ocbo_segment_argmax(const double* F, int S, int ldf, int64_t s_f, const int32_t* seg_ptr,
                        int nseg, double* out_max, int32_t* out_arg, int batch,
                        ocbo_stream_t stream) {
  if (!F) return fail_arg(1, "F null");
  if (S <= 0) return fail_arg(2, "S");
  if (!seg_ptr) return fail_arg(5, "seg_ptr null");
  if (nseg <= 0) return fail_arg(6, "nseg");
  if (!out_max || !out_arg) return fail_arg(7, "outputs null");
  if (batch <= 0) return fail_arg(9, "batch");
  CK(launch_segment_argmax(F, S, ldf, s_f, seg_ptr, nseg, out_max, out_arg, batch,
                           (cudaStream_t)stream),
     "ocbo_segment_argmax");
  return 0;
}

int ocbo_joint_pick(const double* seg_max, const int32_t* seg_arg, int nseg, int n_old_seg, int T,
                    int S, int32_t* out_task, int32_t* out_action, double* out_gain, int batch,
                    ocbo_stream_t stream) {
  if (!seg_max || !seg_arg) return fail_arg(1, "inputs null");
  if (T <= 0 || n_old_seg + T > nseg) return fail_arg(5, "T / n_old_seg inconsistent with nseg");
  if (n_old_seg != 0 && n_old_seg != T) return fail_arg(4, "n_old_seg must be 0 or T");
  if (S <= 0) return fail_arg(6, "S");
  if (!out_task || !out_action || !out_gain) return fail_arg(7, "outputs null");
  if (batch <= 0) return fail_arg(10, "batch");
  CK(launch_joint_pick(seg_max, seg_arg, nseg, n_old_seg, T, S, out_task, out_action, out_gain,
                       batch, (cudaStream_t)stream),
     "ocbo_joint_pick");
  return 0;
}

int ocbo_kg(const double* judge_means, int ld_means, const double* inter, int64_t ld_inter,
            const double* cand_var, double noise, int C, int G, int E, double* kg,
            ocbo_stream_t stream) {
  if (!judge_means) return fail_arg(1, "judge_means null");
  if (!inter) return fail_arg(3, "inter null");
  if (!cand_var) return fail_arg(5, "cand_var null");
  if (!(noise >= 0.0)) return fail_arg(6, "noise must be >= 0");
  if (C <= 0 || C > 65535) return fail_arg(7, "C must be in [1, 65535]");
  if (G <= 0) return fail_arg(8, "G");
  if (E < 2 || E > 1024) return fail_arg(9, "E must be in [2, 1024]");
  if (ld_means < E) return fail_arg(2, "ld_means < E");
  if (ld_inter < (int64_t)G * E) return fail_arg(4, "ld_inter < G * E");
  if (!kg) return fail_arg(10, "kg null");
  CK(launch_kg(judge_means, ld_means, inter, ld_inter, cand_var, noise, C, G, E, kg,
               (cudaStream_t)stream),
     "ocbo_kg");
  return 0;
}

int ocbo_probe_dmma(double* sink, int blocks, int iters, ocbo_stream_t stream) {
  if (!sink) return fail_arg(1, "sink null");
  CK(launch_probe_dmma(sink, blocks, iters, (cudaStream_t)stream), "ocbo_probe_dmma");
  return 0;
}
int ocbo_probe_dfma(double* sink, int blocks, int iters, ocbo_stream_t stream) {
  if (!sink) return fail_arg(1, "sink null");
  CK(launch_probe_dfma(sink, blocks, iters, (cudaStream_t)stream), "ocbo_probe_dfma");
  return 0;
}
int ocbo_probe_kern_se(const double* d2, double scale, int64_t n, double* lockstep, double* plain,
                       ocbo_stream_t stream) {
  if (n < 0) return fail_arg(3, "n < 0");
  if (n > 0 && (!d2 || !lockstep || !plain)) return fail_arg(1, "null buffer");
  CK(launch_probe_kern_se(d2, scale, n, lockstep, plain, (cudaStream_t)stream), "ocbo_probe_kern_se");
  return 0;
}
int ocbo_probe_igemm_nt(const int8_t* A, int64_t lda, const int8_t* B, int64_t ldb, int32_t* C, int64_t ldc, int m,
                        int n, int k, ocbo_stream_t stream) {
  if (!A || !B || !C) return fail_arg(1, "null");
  if (m <= 0 || m % 128 || n <= 0 || n % 128) return fail_arg(7, "m and n must be positive multiples of 128");
  if (k <= 0 || k % 128) return fail_arg(9, "k must be a positive multiple of 128");
  if (lda < k || (lda & 15) || ldb < k || (ldb & 15)) return fail_arg(2, "lda / ldb must be multiples of 16 and >= k");
  if (ldc < n || (ldc & 3)) return fail_arg(6, "ldc must be a multiple of 4 and >= n");
  CK(launch_probe_igemm(A, lda, B, ldb, C, ldc, m, n, k, (cudaStream_t)stream), "ocbo_probe_igemm_nt");
  return 0;
}
int ocbo_ozaki_slice_pairs(void) { return ozaki_slice_pairs(); }
int64_t ocbo_ozaki_ws_bytes(int rows, int kw, int batch) {
  if (rows <= 0 || kw <= 0 || batch <= 0) return 0;
  return (int64_t)ozaki_ws_bytes(rows, kw, batch);
}
int ocbo_probe_ozaki_update(const double* P, int ldp, int rows, int kw, double* C, int ldc, void* ws, int64_t ws_bytes,
                            int tri, ocbo_stream_t stream) {
  if (!P || !C || !ws) return fail_arg(1, "null");
  if (rows <= 0 || rows % 128) return fail_arg(3, "rows must be a positive multiple of 128");
  if (kw <= 0 || kw % 32 || kw > 16384) return fail_arg(4, "kw must be a multiple of 32 in [32, 16384]");
  if (ldp < kw || (ldp & 3)) return fail_arg(2, "ldp must be a multiple of 4 and >= kw");
  if ((uintptr_t)P & 31) return fail_arg(1, "P must be 32-byte aligned");
  if ((uintptr_t)ws & 255) return fail_arg(7, "workspace must be 256-byte aligned");
  if (ldc < rows || (ldc & 1)) return fail_arg(6, "ldc");
  if (ws_bytes < (int64_t)ozaki_ws_bytes(rows, kw, 1)) return fail_arg(8, "workspace too small (ocbo_ozaki_ws_bytes)");
  CK(launch_ozaki_split(P, ldp, 0, rows, kw, ws, nullptr, 1, (cudaStream_t)stream), "ozaki split");
  CK(launch_ozaki_update(ws, rows, kw, 0, 0, C, ldc, 0, rows / 128, rows / 64, tri, 0, 0, 0, nullptr, 1,
                         (cudaStream_t)stream),
     "ozaki update");
  return 0;
}
int ocbo_probe_dgemm_nt(const double* A, const double* B, double* C, int n, ocbo_stream_t stream) {
  if (!A || !B || !C) return fail_arg(1, "null");
  if (!tile_ok(n)) return fail_arg(4, "n must be a multiple of 128");
  NtParams p{};
  p.P = A; p.ldp = n; p.sP = 0;
  p.Q = B; p.ldq = n; p.sQ = 0;
  p.C = C; p.ldc = n; p.sC = 0;
  p.ntc = n / TN; p.ktiles = n / BK; p.tri = 0; p.lower_skip = 0; p.mode = 0; p.info = nullptr;
  CK(launch_gemm_nt(p, (n / TILE) * (n / TN), 1, (cudaStream_t)stream), "ocbo_probe_dgemm_nt");
  return 0;
}

}  // extern "C"
