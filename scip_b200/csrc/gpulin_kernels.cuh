// gpulin_kernels.cuh -- the propagation round of the linear bound propagation path as sm_100a kernels.
//
// One round (= one sweep of the reference's consPropLinear over its marked rows, cons_linear.c:16126-16195, run as a
// synchronous Jacobi step) is two launches:
//
//   sweep_rows_kernel / sweep_long_kernel     for every row marked for propagation: min/max activity with
//        inf/huge counters (double-double), row gates, per-nonzero candidate bounds, atomicMin on int64 keys,
//        row verdict.  Rows are binned by length: thread-per-row on SELL-32 slices (len <= 32), warp-per-row
//        (33..1024), block-per-row with shared-memory staging (> 1024).  The matrix is streamed ONCE per round:
//        between the activity pass and the candidate pass only alpha_k = |a_k| (ub_k - lb_k) stays in registers /
//        shared memory; a nonzero is re-read only if its alpha passes the slack test (tightenVarBoundsEasy
//        :5474/:5566), which is rare.
//   apply_kernel                              for every column whose key moved: accept the new bounds, detect
//        crossing bounds, log the change, mark the rows of the column (CSC) for the next round -- the counterpart
//        of eventExecLinear's SCIPmarkConsPropagate (:17229); the last block to finish runs the loop control
//        (propagateDomains, solve.c:766) and sets the CUDA-graph WHILE condition.
#pragma once

#include "gpulin_device.cuh"

namespace gpl {

constexpr int SWEEP_THREADS = 256;
constexpr int LONG_THREADS = 512;
constexpr int MAX_HIST = 1024;       // rounds with recorded per-round statistics
constexpr int SHORT_MAXLEN = 32;     // thread-per-row up to this length
constexpr int MEDIUM_MAXLEN = 1024;  // warp-per-row up to this length
constexpr int NNZ_SLOTS = 64;        // spread counters of the nonzeros swept in a round

// loop control + statistics, lives in device memory
struct Ctrl
{
   int                maxrounds;    // <= 0: unlimited            } written by the host before every call
   int                logcap;       // capacity of the change log } (one 8-byte copy from pinned memory)
   int                round;
   int                cont;         // 1: another round follows
   int                status;       // GPULIN_FIXPOINT / _CUTOFF / _ROUNDLIMIT
   int                cutoff;       // set by any kernel that proves infeasibility
   unsigned int       ticket;       // apply kernel: blocks finished
   unsigned int       pad0;
   unsigned long long logcount;     // entries produced
   unsigned long long round_nchg;   // accepted bound changes of the running round
   unsigned long long total_nchg;
   unsigned long long total_nnz;
   unsigned long long t_start;      // %globaltimer at the start of the call
   unsigned int       nexact[4];    // rows the filter sweeps handed to the exact kernel in the running round, per bin
   unsigned long long round_nnz[NNZ_SLOTS];  // nonzeros swept in the running round (sum over the slots)
   unsigned long long hist_time[MAX_HIST];   // %globaltimer at the end of each round
   unsigned long long hist_nnz[MAX_HIST];
   unsigned long long hist_nchg[MAX_HIST];
};

struct ChangeRec      // == gpulin_change
{
   int    var;
   int    round;
   double newbound;
   int    is_upper;
   int    reserved;
};

struct DevProblem
{
   int                 nrows;
   int                 ncols;
   int                 nshort;     // rows [0,nshort) are short (SELL-32), then medium, then long (CSR)
   int                 nmedium;
   // rows in permuted order
   const long long*    sell_off;   // per slice of 32 short rows: element offset of the slice
   const int*          rowlen;     // per row
   const long long*    rowbeg;     // per row: element offset of its first nonzero (CSR part; unused for short rows)
   const double*       vals;
   const int*          cols;       // column index | (integral << 31)
   const double2*      sides;      // (lhs, rhs) per row
   unsigned char*      dirty;      // per row: marked for propagation
   int*                xlist;      // rows handed to exact_rows_kernel; one region per bin: [0,nshort) [nshort,+nmedium) [..,nrows)
   // columns
   const double2*      bnd;        // (lb, ub) at round start
   long long*          cand;       // 2*ncols (+2) candidate keys, see Sink
   unsigned char*      colflag;
   // column -> rows (permuted row ids)
   const long long*    colbeg;
   const int*          colrows;
   Ctrl*               ctrl;
   ChangeRec*          log;
   Num                 num;
};

__device__ __forceinline__ unsigned long long globaltimer()
{
   unsigned long long t;
   asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
   return t;
}

// streaming loads of the matrix: read once per round, evict-first
__device__ __forceinline__ double ldStream(const double* p) { return __ldcs(p); }
__device__ __forceinline__ int ldStream(const int* p) { return __ldcs(p); }

// ---- butterfly reduction of the row state over the lanes of a warp; both partners of an exchange merge
// ---- (lower lane, upper lane) in the same order, so every lane ends with the identical double-double value
__device__ __forceinline__ void accWarpReduce(RowAcc& r, int lane)
{
#pragma unroll
   for( int m = 16; m >= 1; m >>= 1 )
   {
      RowAcc o;
      o.minhi = __shfl_xor_sync(0xffffffffu, r.minhi, m);
      o.minlo = __shfl_xor_sync(0xffffffffu, r.minlo, m);
      o.maxhi = __shfl_xor_sync(0xffffffffu, r.maxhi, m);
      o.maxlo = __shfl_xor_sync(0xffffffffu, r.maxlo, m);
      o.maxdelta = __shfl_xor_sync(0xffffffffu, r.maxdelta, m);
      o.cnt = __shfl_xor_sync(0xffffffffu, r.cnt, m);
      const bool upper = (lane & m) != 0;
      RowAcc a = upper ? o : r;
      const RowAcc b = upper ? r : o;
      accMerge(a, b);
      r = a;
   }
}

__device__ __forceinline__ bool passesSlackTest(const RowInfo& ri, double alpha, double thr)
{
   return (ri.rhsfin && alpha - ri.slackR > thr) || (ri.lhsfin && alpha - ri.slackL > thr);
}

// threshold of the slack test: alpha - slack > sumepsilon, or > epsilon for single-variable rows (:5474, :5566)
__device__ __forceinline__ double slackThreshold(const Num& n, bool force)
{
   return force ? fmin(n.eps, n.sumeps) : n.sumeps;
}

// ---- candidate pass over the elements first, first+step, ... < len of a row that passed the gates of tightenBounds
// ---- (rare once the bounds have settled): the row is read a second time, from L2.  Easy rows only look at nonzeros
// ---- whose alpha = |a| (ub - lb) exceeds the slack (tightenVarBoundsEasy :5474/:5566).
__device__ __forceinline__ void rowCandidates(const DevProblem& p, const RowInfo& ri, long long base, int stride,
   int first, int step, int len, bool& cutoff)
{
   const Num& n = p.num;
   const double thr = slackThreshold(n, ri.force);
   Sink s;
   s.cand = p.cand;
   s.colflag = p.colflag;
   for( int k0 = first; k0 < len; k0 += 4 * step )
   {
      double a[4];
      int cj[4];
      double2 b[4];
#pragma unroll
      for( int q = 0; q < 4; ++q )
      {
         const int k = k0 + q * step;
         if( k < len )
         {
            a[q] = p.vals[base + (long long)stride * k];
            cj[q] = p.cols[base + (long long)stride * k];
         }
      }
#pragma unroll
      for( int q = 0; q < 4; ++q )
      {
         if( k0 + q * step < len )
            b[q] = p.bnd[cj[q] & 0x7fffffff];
      }
#pragma unroll
      for( int q = 0; q < 4; ++q )
      {
         if( k0 + q * step < len )
         {
            if( !ri.easy || passesSlackTest(ri, fabs(a[q]) * (b[q].y - b[q].x), thr) )
               candidates(n, s, ri, a[q], cj[q] & 0x7fffffff, cj[q] < 0, b[q].x, b[q].y, cutoff);
         }
      }
   }
}

// ---- hot path: an interval filter in plain fp64 ----------------------------------------------------------------
// The sweep accumulates, per row, the activities minact/maxact as ordinary double sums together with
//    maxdelta ~ max_k |a_k| (ub_k - lb_k)      and      cabs = max_k max(|a_k lb_k|, |a_k ub_k|),
// 9 fp64 instructions per nonzero.  With n = row length and u = 2^-52 every one of these differs from the value the
// exact rules would use (double-double sums, gpulin_device.cuh) by at most E = (n^2 + 8) u cabs.  A row is CLEARLY
// QUIET if, for every perturbation within E, the reference's gates say "no bound can be tightened"
// (tightenBounds :7057, :7081) and its verdict says "feasible" (propagateCons :7728).  Such a row is finished --
// with exactly the result of the full rules.  Every other row (it may tighten something, it may be infeasible, it
// holds an infinite bound or a huge product, or it is too close to call) is redone by rowExact*: double-double
// activities with the inf/huge counters, gates, candidates, verdict, as restated from the reference.
// Once the bounds have settled almost all rows are clearly quiet, so the matrix is streamed once per round and the
// exact pass touches (from L2) only rows that have work.
constexpr int ROWLEN_EXACT = 0x40000000;   // flag in rowlen[]: the row has a coefficient so small that an infinite
                                           // bound could hide behind a non-huge product -> always the exact rules

struct LeanAcc
{
   double minact, maxact, maxdelta, cabs;
};

__device__ __forceinline__ void leanInit(LeanAcc& r)
{
   r.minact = r.maxact = r.maxdelta = r.cabs = 0.0;
}

__device__ __forceinline__ void leanElem(LeanAcc& r, double a, double l, double u)
{
   const double al = a * l;
   const double au = a * u;
   const double cmin = fmin(al, au);
   const double cmax = fmax(al, au);
   r.minact += cmin;
   r.maxact += cmax;
   r.maxdelta = fmax(r.maxdelta, cmax - cmin);
   r.cabs = fmax(r.cabs, fmax(fabs(al), fabs(au)));
}

__device__ __forceinline__ void leanWarpReduce(LeanAcc& r)
{
#pragma unroll
   for( int m = 16; m >= 1; m >>= 1 )
   {
      r.minact += __shfl_xor_sync(0xffffffffu, r.minact, m);
      r.maxact += __shfl_xor_sync(0xffffffffu, r.maxact, m);
      r.maxdelta = fmax(r.maxdelta, __shfl_xor_sync(0xffffffffu, r.maxdelta, m));
      r.cabs = fmax(r.cabs, __shfl_xor_sync(0xffffffffu, r.cabs, m));
   }
}

__device__ __forceinline__ bool rowClearlyQuiet(const Num& n, const LeanAcc& r, int len, double lhs, double rhs)
{
   // an infinite bound (|b| >= 1e20, with |a| >= hugeval/infinity by the ROWLEN_EXACT flag) or a huge product shows
   // up in cabs; NaN compares false
   if( !(r.cabs < n.huge) )
      return false;
   const double E = (double)(len * len + 8) * 2.3e-16 * r.cabs;
   // verdict: FeasGT(minact,rhs) needs (minact-rhs)/max(1,|minact|,|rhs|) > feastol, impossible if minact-rhs <= feastol/4
   if( (r.minact - rhs) + E > 0.25 * n.feastol || (lhs - r.maxact) + E > 0.25 * n.feastol )
      return false;
   const double md = r.maxdelta + E;
   if( md <= 0.5 * n.feastol )
      return true;                // all variables fixed (:7057)
   const double slack = isInf(n, rhs) ? n.inf : rhs - r.minact;
   const double surplus = isInf(n, -lhs) ? n.inf : r.maxact - lhs;
   const double m = fmin(slack, surplus);
   // the gate maxdelta <= min(slack, surplus) + eps (:7081) with every rounding in its disfavour
   return md + 2.0 * E + 4.5e-16 * fabs(m) - m <= 0.5 * n.eps;
}

// gates, candidate pass and verdict of one row whose exact activities are known (elements first, first+step, ... of
// the calling thread)
__device__ __forceinline__ void rowTighten(const DevProblem& p, const RowAcc& acc, double lhs, double rhs, long long base,
   int stride, int first, int step, int len)
{
   RowInfo ri;
   ri.acc = acc;
   ri.lhs = lhs;
   ri.rhs = rhs;
   bool cutoff = false;
   if( rowGates(p.num, ri, len, cutoff) )
      rowCandidates(p, ri, base, stride, first, step, len, cutoff);
   if( cutoff || rowInfeasible(p.num, ri.acc, lhs, rhs) )
      p.ctrl->cutoff = 1;
}

// exact activities of the elements first, first+step, ... < len (four independent loads in flight)
__device__ __forceinline__ void accumulateExact(const DevProblem& p, RowAcc& acc, long long base, int stride, int first,
   int step, int len)
{
   for( int k0 = first; k0 < len; k0 += 4 * step )
   {
      double a[4];
      int cj[4];
      double2 b[4];
#pragma unroll
      for( int q = 0; q < 4; ++q )
      {
         const int k = k0 + q * step;
         if( k < len )
         {
            a[q] = p.vals[base + (long long)stride * k];
            cj[q] = p.cols[base + (long long)stride * k];
         }
      }
#pragma unroll
      for( int q = 0; q < 4; ++q )
      {
         if( k0 + q * step < len )
            b[q] = p.bnd[cj[q] & 0x7fffffff];
      }
#pragma unroll
      for( int q = 0; q < 4; ++q )
      {
         if( k0 + q * step < len )
            accElem(p.num, acc, a[q], b[q].x, b[q].y);
      }
   }
}

constexpr unsigned char ROW_CLEAN = 0;
constexpr unsigned char ROW_MARKED = 1;

// butterfly over the W lanes of a sub-warp group (all 32 lanes of the warp execute it)
template <int W>
__device__ __forceinline__ void accGroupReduce(RowAcc& r, int lane)
{
#pragma unroll
   for( int m = W / 2; m >= 1; m >>= 1 )
   {
      RowAcc o;
      o.minhi = __shfl_xor_sync(0xffffffffu, r.minhi, m);
      o.minlo = __shfl_xor_sync(0xffffffffu, r.minlo, m);
      o.maxhi = __shfl_xor_sync(0xffffffffu, r.maxhi, m);
      o.maxlo = __shfl_xor_sync(0xffffffffu, r.maxlo, m);
      o.maxdelta = __shfl_xor_sync(0xffffffffu, r.maxdelta, m);
      o.cnt = __shfl_xor_sync(0xffffffffu, r.cnt, m);
      const bool upper = (lane & m) != 0;
      RowAcc a = upper ? o : r;
      const RowAcc b = upper ? r : o;
      accMerge(a, b);
      r = a;
   }
}

__device__ __forceinline__ void addRoundNnz(const DevProblem& p, unsigned long long nnzdone, int slot)
{
   if( nnzdone != 0 )
      atomicAdd(&p.ctrl->round_nnz[slot & (NNZ_SLOTS - 1)], nnzdone);
}

// ---- thread-per-row on SELL-32 slices: element k of the row of lane t sits at slice_off + 32 k + t -------------
// ---- persistent warps: warp w sweeps slices w, w + W, w + 2W, ...
template <int CH>
__device__ __forceinline__ void loadChunk(const DevProblem& p, long long base, int c, int len, double (&a)[CH], int (&cj)[CH])
{
#pragma unroll
   for( int k = 0; k < CH; ++k )
   {
      if( c + k < len )
      {
         a[k] = ldStream(p.vals + base + 32LL * (c + k));
         cj[k] = ldStream(p.cols + base + 32LL * (c + k));
      }
   }
}

template <int CH, bool PF>
__device__ __forceinline__ void sweepSlice(const DevProblem& p, int slice, int lane, bool act, unsigned& nnzdone)
{
   const Num& n = p.num;
   const int row = slice * 32 + lane;
   int len = 0;
   bool exact = false;
   if( act )
   {
      len = p.rowlen[row];
      exact = (len & ROWLEN_EXACT) != 0;
      len &= ~ROWLEN_EXACT;
   }
   const long long base = p.sell_off[slice] + lane;
   const int maxlen = __reduce_max_sync(0xffffffffu, len);

   LeanAcc acc;
   leanInit(acc);
   if( PF )
   {
      // the coefficients of chunk c+1 are in flight while the bounds of chunk c are gathered
      double an[CH];
      int cjn[CH];
      loadChunk<CH>(p, base, 0, len, an, cjn);
      for( int c = 0; c < maxlen; c += CH )
      {
         double a[CH];
         int cj[CH];
         double2 b[CH];
#pragma unroll
         for( int k = 0; k < CH; ++k )
         {
            a[k] = an[k];
            cj[k] = cjn[k];
         }
         if( c + CH < maxlen )
            loadChunk<CH>(p, base, c + CH, len, an, cjn);
#pragma unroll
         for( int k = 0; k < CH; ++k )
         {
            if( c + k < len )
               b[k] = p.bnd[cj[k] & 0x7fffffff];
         }
#pragma unroll
         for( int k = 0; k < CH; ++k )
         {
            if( c + k < len )
               leanElem(acc, a[k], b[k].x, b[k].y);
         }
      }
   }
   else
   {
      for( int c = 0; c < maxlen; c += CH )
      {
         double a[CH];
         int cj[CH];
         double2 b[CH];
         loadChunk<CH>(p, base, c, len, a, cj);
#pragma unroll
         for( int k = 0; k < CH; ++k )
         {
            if( c + k < len )
               b[k] = p.bnd[cj[k] & 0x7fffffff];
         }
#pragma unroll
         for( int k = 0; k < CH; ++k )
         {
            if( c + k < len )
               leanElem(acc, a[k], b[k].x, b[k].y);
         }
      }
   }
   bool handoff = false;
   if( act )
   {
      const double2 sd = p.sides[row];
      handoff = exact || !rowClearlyQuiet(n, acc, len, sd.x, sd.y);
      p.dirty[row] = ROW_CLEAN;
      nnzdone += (unsigned)len;
   }
   // rows that need the exact rules go to the work list of exact_rows_kernel (one atomic per warp)
   const unsigned hm = __ballot_sync(0xffffffffu, handoff);
   if( hm != 0u )
   {
      unsigned pos = 0u;
      if( lane == 0 )
         pos = atomicAdd(&p.ctrl->nexact[0], (unsigned)__popc(hm));
      pos = __shfl_sync(0xffffffffu, pos, 0);
      if( handoff )
         p.xlist[pos + __popc(hm & ((1u << lane) - 1u))] = row;
   }
}

template <int CH, bool PF, int MINB>
__global__ void __launch_bounds__(SWEEP_THREADS, MINB) sweep_short_kernel(const DevProblem p)
{
   const int lane = threadIdx.x & 31;
   const int gw = (blockIdx.x * SWEEP_THREADS + threadIdx.x) >> 5;
   const int nw = (gridDim.x * SWEEP_THREADS) >> 5;
   const int nslices = (p.nshort + 31) >> 5;
   unsigned nnzdone = 0;
   for( int s0 = gw; s0 < nslices; s0 += 4 * nw )
   {
      // the flags of the next four slices of this warp, one load latency
      unsigned fm = 0u;
#pragma unroll
      for( int i = 0; i < 4; ++i )
      {
         const int row = (s0 + i * nw) * 32 + lane;
         if( s0 + i * nw < nslices && row < p.nshort && p.dirty[row] == ROW_MARKED )
            fm |= 1u << i;
      }
#pragma unroll 1
      for( int i = 0; i < 4; ++i )
      {
         const bool act = ((fm >> i) & 1u) != 0u;
         if( __any_sync(0xffffffffu, act) )
            sweepSlice<CH, PF>(p, s0 + i * nw, lane, act, nnzdone);
      }
   }
   nnzdone = __reduce_add_sync(0xffffffffu, nnzdone);
   if( lane == 0 )
      addRoundNnz(p, (unsigned long long)nnzdone, gw);
}

// ---- warp-per-row on CSR (rows of 33..1024 nonzeros): lane t reads elements t, t+32, ... ----------------------
template <int U>
__device__ __forceinline__ void loadChunkW(const DevProblem& p, long long beg, int c, int lane, int len, double (&a)[U], int (&cj)[U])
{
#pragma unroll
   for( int k = 0; k < U; ++k )
   {
      const int idx = c + k * 32 + lane;
      if( idx < len )
      {
         a[k] = ldStream(p.vals + beg + idx);
         cj[k] = ldStream(p.cols + beg + idx);
      }
   }
}

template <int U>
__device__ __forceinline__ void sweepWarpRow(const DevProblem& p, int row, int lane, unsigned long long& nnzdone)
{
   const Num& n = p.num;
   int len = p.rowlen[row];
   const bool exact = (len & ROWLEN_EXACT) != 0;
   len &= ~ROWLEN_EXACT;
   const long long beg = p.rowbeg[row];
   const double2 sd = p.sides[row];

   LeanAcc acc;
   leanInit(acc);
   double an[U];
   int cjn[U];
   loadChunkW<U>(p, beg, 0, lane, len, an, cjn);
   for( int c = 0; c < len; c += 32 * U )
   {
      double a[U];
      int cj[U];
      double2 b[U];
#pragma unroll
      for( int k = 0; k < U; ++k )
      {
         a[k] = an[k];
         cj[k] = cjn[k];
      }
      if( c + 32 * U < len )
         loadChunkW<U>(p, beg, c + 32 * U, lane, len, an, cjn);
#pragma unroll
      for( int k = 0; k < U; ++k )
      {
         if( c + k * 32 + lane < len )
            b[k] = p.bnd[cj[k] & 0x7fffffff];
      }
#pragma unroll
      for( int k = 0; k < U; ++k )
      {
         if( c + k * 32 + lane < len )
            leanElem(acc, a[k], b[k].x, b[k].y);
      }
   }
   leanWarpReduce(acc);
   const bool handoff = exact || !rowClearlyQuiet(n, acc, len, sd.x, sd.y);   // warp-uniform: all lanes hold the same sums
   if( lane == 0 )
   {
      p.dirty[row] = ROW_CLEAN;
      nnzdone += (unsigned long long)len;
      if( handoff )
         p.xlist[p.nshort + atomicAdd(&p.ctrl->nexact[1], 1u)] = row;
   }
}

template <int U, int MINB>
__global__ void __launch_bounds__(SWEEP_THREADS, MINB) sweep_medium_kernel(const DevProblem p)
{
   const int lane = threadIdx.x & 31;
   const int gw = (blockIdx.x * SWEEP_THREADS + threadIdx.x) >> 5;
   const int nw = (gridDim.x * SWEEP_THREADS) >> 5;
   const int row0 = p.nshort;
   const int nrows = p.nmedium;
   unsigned long long nnzdone = 0;
   for( int r0 = gw; r0 < nrows; r0 += 32 * nw )
   {
      // lane i looks at the flag of the i-th next row of this warp
      const int mine = r0 + lane * nw;
      const bool f = (mine < nrows) && p.dirty[row0 + mine] == ROW_MARKED;
      unsigned mask = __ballot_sync(0xffffffffu, f);
      while( mask != 0u )
      {
         const int i = __ffs(mask) - 1;
         mask &= mask - 1u;
         sweepWarpRow<U>(p, row0 + r0 + i * nw, lane, nnzdone);
      }
   }
   if( lane == 0 )
      addRoundNnz(p, nnzdone, gw);
}

// ---- block-per-row for rows longer than MEDIUM_MAXLEN -----------------------------------------------------------
__global__ void __launch_bounds__(LONG_THREADS) sweep_long_kernel(const DevProblem p)
{
   __shared__ LeanAcc s_lean[LONG_THREADS / 32];

   const int lane = threadIdx.x & 31;
   const int warp = threadIdx.x >> 5;
   const Num& n = p.num;
   const int row0 = p.nshort + p.nmedium;
   const int nrows = p.nrows - row0;

   for( int r = blockIdx.x; r < nrows; r += gridDim.x )
   {
      const int row = row0 + r;
      __syncthreads();                  // previous row done with the shared state and its flag
      if( p.dirty[row] != ROW_MARKED )  // block-uniform: nobody rewrites the flag before the barriers below
         continue;
      int len = p.rowlen[row];
      const bool exact = (len & ROWLEN_EXACT) != 0;
      len &= ~ROWLEN_EXACT;
      const long long beg = p.rowbeg[row];
      const double2 sd = p.sides[row];

      LeanAcc la;
      leanInit(la);
      for( int i0 = 0; i0 < len; i0 += 4 * LONG_THREADS )
      {
         double a[4];
         int cj[4];
         double2 b[4];
#pragma unroll
         for( int k = 0; k < 4; ++k )
         {
            const int idx = i0 + k * LONG_THREADS + threadIdx.x;
            if( idx < len )
            {
               a[k] = ldStream(p.vals + beg + idx);
               cj[k] = ldStream(p.cols + beg + idx);
            }
         }
#pragma unroll
         for( int k = 0; k < 4; ++k )
         {
            if( i0 + k * LONG_THREADS + threadIdx.x < len )
               b[k] = p.bnd[cj[k] & 0x7fffffff];
         }
#pragma unroll
         for( int k = 0; k < 4; ++k )
         {
            if( i0 + k * LONG_THREADS + threadIdx.x < len )
               leanElem(la, a[k], b[k].x, b[k].y);
         }
      }
      leanWarpReduce(la);
      if( lane == 0 )
         s_lean[warp] = la;
      __syncthreads();
      if( threadIdx.x == 0 )
      {
         la = s_lean[0];
#pragma unroll 1
         for( int w = 1; w < LONG_THREADS / 32; ++w )
         {
            la.minact += s_lean[w].minact;
            la.maxact += s_lean[w].maxact;
            la.maxdelta = fmax(la.maxdelta, s_lean[w].maxdelta);
            la.cabs = fmax(la.cabs, s_lean[w].cabs);
         }
         const bool handoff = exact || !rowClearlyQuiet(n, la, len, sd.x, sd.y);
         p.dirty[row] = ROW_CLEAN;
         addRoundNnz(p, (unsigned long long)len, row);
         if( handoff )
            p.xlist[p.nshort + p.nmedium + atomicAdd(&p.ctrl->nexact[2], 1u)] = row;
      }
   }
}

// ---- the exact rules for every row the filter sweeps could not finish ---------------------------------------------
// one launch over the three work lists: short rows by groups of 8 lanes (the whole row in registers: activities by a
// butterfly over the group, candidates straight from the registers), medium rows by warps, long rows by blocks
constexpr int EXACT_THREADS = 256;
constexpr int EXACT_G = 8;                           // lanes per short row
constexpr int EXACT_Q = SHORT_MAXLEN / EXACT_G;      // elements per lane

__global__ void __launch_bounds__(EXACT_THREADS) exact_rows_kernel(const DevProblem p)
{
   __shared__ RowAcc s_acc[EXACT_THREADS / 32];

   const unsigned n0 = p.ctrl->nexact[0];
   const unsigned n1 = p.ctrl->nexact[1];
   const unsigned n2 = p.ctrl->nexact[2];
   if( (n0 | n1 | n2) == 0u )
      return;
   const Num& n = p.num;
   const int lane = threadIdx.x & 31;
   const int warp = threadIdx.x >> 5;
   const int gtid = blockIdx.x * EXACT_THREADS + threadIdx.x;
   const int nthreads = gridDim.x * EXACT_THREADS;

   // ---- short rows: EXACT_G lanes per row
   {
      const int gl = lane & (EXACT_G - 1);
      const int ngroups = nthreads / EXACT_G;
      const unsigned rounds = (n0 + ngroups - 1) / ngroups;      // warp-uniform trip count
      for( unsigned it = 0; it < rounds; ++it )
      {
         const unsigned item = it * ngroups + gtid / EXACT_G;
         const bool valid = item < n0;
         int len = 0;
         long long base = 0;
         double2 sd = make_double2(0.0, 0.0);
         if( valid )
         {
            const int row = p.xlist[item];
            len = p.rowlen[row] & ~ROWLEN_EXACT;
            base = p.sell_off[row >> 5] + (row & 31);
            sd = p.sides[row];
         }
         double a[EXACT_Q];
         int cj[EXACT_Q];
         double2 b[EXACT_Q];
#pragma unroll
         for( int q = 0; q < EXACT_Q; ++q )
         {
            const int k = gl + EXACT_G * q;
            if( k < len )
            {
               a[q] = p.vals[base + 32LL * k];
               cj[q] = p.cols[base + 32LL * k];
            }
         }
#pragma unroll
         for( int q = 0; q < EXACT_Q; ++q )
         {
            if( gl + EXACT_G * q < len )
               b[q] = p.bnd[cj[q] & 0x7fffffff];
         }
         RowInfo ri;
         accInit(ri.acc);
#pragma unroll
         for( int q = 0; q < EXACT_Q; ++q )
         {
            if( gl + EXACT_G * q < len )
               accElem(n, ri.acc, a[q], b[q].x, b[q].y);
         }
         accGroupReduce<EXACT_G>(ri.acc, lane);
         if( valid )
         {
            ri.lhs = sd.x;
            ri.rhs = sd.y;
            bool cutoff = false;
            if( rowGates(n, ri, len, cutoff) )      // uniform over the group
            {
               const double thr = slackThreshold(n, ri.force);
               Sink sk;
               sk.cand = p.cand;
               sk.colflag = p.colflag;
#pragma unroll
               for( int q = 0; q < EXACT_Q; ++q )
               {
                  if( gl + EXACT_G * q < len )
                  {
                     if( !ri.easy || passesSlackTest(ri, fabs(a[q]) * (b[q].y - b[q].x), thr) )
                        candidates(n, sk, ri, a[q], cj[q] & 0x7fffffff, cj[q] < 0, b[q].x, b[q].y, cutoff);
                  }
               }
            }
            if( cutoff || (gl == 0 && rowInfeasible(n, ri.acc, ri.lhs, ri.rhs)) )
               p.ctrl->cutoff = 1;
         }
      }
   }

   // ---- medium rows: one warp per row
   {
      const int gw = gtid >> 5;
      const int nw = nthreads >> 5;
      for( unsigned item = gw; item < n1; item += nw )
      {
         const int row = p.xlist[p.nshort + item];
         const int len = p.rowlen[row] & ~ROWLEN_EXACT;
         const long long beg = p.rowbeg[row];
         const double2 sd = p.sides[row];
         RowAcc acc;
         accInit(acc);
         accumulateExact(p, acc, beg, 1, lane, 32, len);
         accWarpReduce(acc, lane);
         rowTighten(p, acc, sd.x, sd.y, beg, 1, lane, 32, len);
      }
   }

   // ---- long rows: one block per row
   for( unsigned item = blockIdx.x; item < n2; item += gridDim.x )
   {
      const int row = p.xlist[p.nshort + p.nmedium + item];
      const int len = p.rowlen[row] & ~ROWLEN_EXACT;
      const long long beg = p.rowbeg[row];
      const double2 sd = p.sides[row];
      RowAcc acc;
      accInit(acc);
      accumulateExact(p, acc, beg, 1, threadIdx.x, EXACT_THREADS, len);
      accWarpReduce(acc, lane);
      __syncthreads();                  // the previous row's s_acc has been read by everybody
      if( lane == 0 )
         s_acc[warp] = acc;
      __syncthreads();
      acc = s_acc[0];
#pragma unroll 1
      for( int w = 1; w < EXACT_THREADS / 32; ++w )
         accMerge(acc, s_acc[w]);
      rowTighten(p, acc, sd.x, sd.y, beg, 1, threadIdx.x, EXACT_THREADS, len);
   }
}

// ---- accept the new bounds of one column; returns the number of changed bounds (0..2) -------------------------
__device__ __forceinline__ int applyColumn(const DevProblem& p, int j, double2& nb, bool& lbchg, bool& ubchg)
{
   const Num& n = p.num;
   const double2 old = p.bnd[j];
   const longlong2 k = reinterpret_cast<const longlong2*>(p.cand)[j];
   double nl = key2d(~k.x);
   double nu = key2d(k.y);
   if( nl > nu )
   {
      // both bounds moved in the same round and crossed: infeasible unless within feastol (scip_var.c:6994/7100)
      if( isFeasGT(n, nl, nu) )
         p.ctrl->cutoff = 1;
      nl = nu;
      p.cand[2 * (size_t)j] = ~d2key(nl);
   }
   lbchg = (nl != old.x);
   ubchg = (nu != old.y);
   nb = make_double2(nl, nu);
   if( lbchg || ubchg )
      const_cast<double2*>(p.bnd)[j] = nb;
   return (int)lbchg + (int)ubchg;
}

__device__ __forceinline__ void markColumnRows(const DevProblem& p, int j)
{
   // row ids are fetched eight at a time before any flag is stored: the byte stores may alias anything as far as the
   // compiler knows, and a load-store-load-store chain would cost one memory round trip per row
   const long long e = p.colbeg[j + 1];
   for( long long q = p.colbeg[j]; q < e; q += 8 )
   {
      int r[8];
#pragma unroll
      for( int t = 0; t < 8; ++t )
      {
         if( q + t < e )
            r[t] = p.colrows[q + t];
      }
#pragma unroll
      for( int t = 0; t < 8; ++t )
      {
         if( q + t < e )
            p.dirty[r[t]] = ROW_MARKED;
      }
   }
}

// loop control (propagateDomains, solve.c:766-787), run by one thread after the last column was applied
template <bool GRAPH>
__device__ __forceinline__ void controlStep(Ctrl* c, cudaGraphConditionalHandle handle)
{
   const unsigned long long nchg = c->round_nchg;
   const int r = c->round;
   unsigned long long nnz = 0;
   for( int i = 0; i < NNZ_SLOTS; ++i )
   {
      nnz += c->round_nnz[i];
      c->round_nnz[i] = 0;
   }
   if( r < MAX_HIST )
   {
      c->hist_time[r] = globaltimer();
      c->hist_nnz[r] = nnz;
      c->hist_nchg[r] = nchg;
   }
   c->total_nchg += nchg;
   c->total_nnz += nnz;
   c->round_nchg = 0;
   c->ticket = 0;
   c->nexact[0] = c->nexact[1] = c->nexact[2] = 0;
   c->round = r + 1;
   int cont = 0;
   if( c->cutoff )
      c->status = 1;
   else if( nchg == 0 )
      c->status = 0;
   else if( c->maxrounds > 0 && r + 1 >= c->maxrounds )
      c->status = 2;
   else
      cont = 1;
   c->cont = cont;
   if( GRAPH )
      cudaGraphSetConditional(handle, (unsigned)cont);
}

// one column whose candidate keys may have moved: accept, log, mark its rows; returns the number of changed bounds
__device__ __forceinline__ int applyAndMark(const DevProblem& p, int j, int round, int logcap)
{
   bool lbchg;
   bool ubchg;
   double2 nb;
   const int nc = applyColumn(p, j, nb, lbchg, ubchg);
   if( nc > 0 )
   {
      markColumnRows(p, j);
      if( logcap > 0 )
      {
         unsigned long long pos = atomicAdd(&p.ctrl->logcount, (unsigned long long)nc);
         if( lbchg )
         {
            if( pos < (unsigned long long)logcap )
            {
               ChangeRec rec;
               rec.var = j; rec.round = round; rec.newbound = nb.x; rec.is_upper = 0; rec.reserved = 0;
               p.log[pos] = rec;
            }
            ++pos;
         }
         if( ubchg && pos < (unsigned long long)logcap )
         {
            ChangeRec rec;
            rec.var = j; rec.round = round; rec.newbound = nb.y; rec.is_upper = 1; rec.reserved = 0;
            p.log[pos] = rec;
         }
      }
   }
   return nc;
}

// DENSE = false: only columns whose flag was raised by the sweep of this round are looked at (single GPU); the flags
//                are scanned 16 per thread;
// DENSE = true : every column compares its (all-reduced) candidate keys with its bounds (rows sharded over ranks:
//                a key may have been moved by another rank)
constexpr int APPLY_THREADS = 256;

template <bool DENSE, bool GRAPH>
__global__ void __launch_bounds__(APPLY_THREADS) apply_kernel(const DevProblem p, cudaGraphConditionalHandle handle)
{
   __shared__ int s_nchg;
   if( threadIdx.x == 0 )
      s_nchg = 0;
   __syncthreads();

   Ctrl* c = p.ctrl;
   const int lane = threadIdx.x & 31;
   const int round = c->round;
   const int logcap = c->logcap;
   int mychg = 0;
   const int stride = gridDim.x * APPLY_THREADS;
   const int gtid = blockIdx.x * APPLY_THREADS + threadIdx.x;
   if( DENSE )
   {
      for( int j = gtid; j < p.ncols; j += stride )
      {
         p.colflag[j] = 0;
         mychg += applyAndMark(p, j, round, logcap);
      }
   }
   else
   {
      const int nvec = (p.ncols + 15) >> 4;     // colflag is allocated and zeroed beyond ncols
      for( int v = gtid; v < nvec; v += stride )
      {
         const uint4 f = reinterpret_cast<const uint4*>(p.colflag)[v];
         if( (f.x | f.y | f.z | f.w) == 0u )
            continue;
         reinterpret_cast<uint4*>(p.colflag)[v] = make_uint4(0u, 0u, 0u, 0u);
         const unsigned w[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
         for( int q = 0; q < 4; ++q )
         {
#pragma unroll
            for( int t = 0; t < 4; ++t )
            {
               if( (w[q] >> (8 * t)) & 0xffu )
                  mychg += applyAndMark(p, 16 * v + 4 * q + t, round, logcap);
            }
         }
      }
   }
   mychg = __reduce_add_sync(0xffffffffu, mychg);
   if( lane == 0 && mychg != 0 )
      atomicAdd(&s_nchg, mychg);
   __syncthreads();
   if( threadIdx.x == 0 )
   {
      if( s_nchg != 0 )
         atomicAdd(&c->round_nchg, (unsigned long long)s_nchg);
      __threadfence();
      const unsigned t = atomicAdd(&c->ticket, 1u);
      if( t == gridDim.x - 1 )
      {
         __threadfence();
         controlStep<GRAPH>(c, handle);
      }
   }
}

// ---- bound (re)initialisation ----------------------------------------------------------------------------------
__global__ void set_bounds_kernel(const DevProblem p, const double* lb, const double* ub)
{
   const int stride = gridDim.x * blockDim.x;
   for( int j = blockIdx.x * blockDim.x + threadIdx.x; j < p.ncols; j += stride )
   {
      const double l = lb[j] + 0.0;
      const double u = ub[j] + 0.0;
      const_cast<double2*>(p.bnd)[j] = make_double2(l, u);
      reinterpret_cast<longlong2*>(p.cand)[j] = make_longlong2(~d2key(l), d2key(u));
      p.colflag[j] = 0;
   }
   for( int r = blockIdx.x * blockDim.x + threadIdx.x; r < p.nrows; r += stride )
      p.dirty[r] = ROW_MARKED;
}

__global__ void update_bounds_kernel(const DevProblem p, long long nupd, const int* idx, const double* lb, const double* ub)
{
   const long long stride = (long long)gridDim.x * blockDim.x;
   for( long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nupd; i += stride )
   {
      const int j = idx[i];
      const double l = lb[i] + 0.0;
      const double u = ub[i] + 0.0;
      const_cast<double2*>(p.bnd)[j] = make_double2(l, u);
      reinterpret_cast<longlong2*>(p.cand)[j] = make_longlong2(~d2key(l), d2key(u));
      p.colflag[j] = 0;
      markColumnRows(p, j);
   }
}

__global__ void get_bounds_kernel(const DevProblem p, double* lb, double* ub)
{
   const int stride = gridDim.x * blockDim.x;
   for( int j = blockIdx.x * blockDim.x + threadIdx.x; j < p.ncols; j += stride )
   {
      const double2 b = p.bnd[j];
      lb[j] = b.x;
      ub[j] = b.y;
   }
}

__global__ void mark_all_kernel(const DevProblem p)
{
   const int stride = gridDim.x * blockDim.x;
   for( int r = blockIdx.x * blockDim.x + threadIdx.x; r < p.nrows; r += stride )
      p.dirty[r] = ROW_MARKED;
}

// start of a gpulin_propagate call: reset the loop state
__global__ void begin_kernel(Ctrl* c)
{
   c->round = 0;
   c->cont = 1;
   c->status = 0;
   c->cutoff = 0;
   c->ticket = 0;
   c->nexact[0] = c->nexact[1] = c->nexact[2] = 0;
   c->logcount = 0;
   c->round_nchg = 0;
   for( int i = 0; i < NNZ_SLOTS; ++i )
      c->round_nnz[i] = 0;
   c->total_nchg = 0;
   c->total_nnz = 0;
   c->t_start = globaltimer();
}

// multi-GPU: the verdict travels in the two spare keys behind the candidate vector (MIN all-reduce)
__global__ void publish_cutoff_kernel(const DevProblem p)
{
   p.cand[2 * (size_t)p.ncols] = p.ctrl->cutoff ? -1LL : 0LL;
   p.cand[2 * (size_t)p.ncols + 1] = 0LL;
}
__global__ void absorb_cutoff_kernel(const DevProblem p)
{
   if( p.cand[2 * (size_t)p.ncols] < 0 )
      p.ctrl->cutoff = 1;
}

} // namespace gpl
