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
   int                pad0;
   unsigned long long logcount;     // entries produced
   unsigned long long round_nchg;   // accepted bound changes of the running round
   unsigned long long total_nchg;
   unsigned long long total_nnz;
   unsigned long long t_start;      // %globaltimer at the start of the call
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
   // columns
   const double2*      bnd;        // (lb, ub) at round start
   long long*          cand;       // 2*ncols (+2) candidate keys, see Sink
   unsigned char*      colflag;
   // column -> rows (permuted row ids)
   const long long*    colbeg;
   const int*          colrows;
   Ctrl*               ctrl;
   ChangeRec*          log;
   const DevProblem*   self;       // copy of this struct in device memory: what the out-of-line (rare path) functions
                                   // read, so that the kernel parameter never needs an address
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

// ---- hot-path row state: the finite parts only.  An element with an infinite bound or a huge product (rare) just
// ---- raises `slow`; such a row is redone by the *Slow functions below with the full classification.
struct FastAcc
{
   double minhi, minlo, maxhi, maxlo, maxdelta;
};

__device__ __forceinline__ void fastInit(FastAcc& r)
{
   r.minhi = r.minlo = r.maxhi = r.maxlo = r.maxdelta = 0.0;
}

// returns false if the element needs the full classification (the sums are garbage then and get discarded)
__device__ __forceinline__ bool fastElem(const Num& n, FastAcc& r, double a, double l, double u)
{
   const bool pos = a > 0.0;
   const double cmin = a * (pos ? l : u);
   const double cmax = a * (pos ? u : l);
   dd_add(r.minhi, r.minlo, cmin);
   dd_add(r.maxhi, r.maxlo, cmax);
   r.maxdelta = fmax(r.maxdelta, fabs(a) * (u - l));
   return (fabs(l) < n.inf) && (fabs(u) < n.inf) && (fabs(cmin) < n.huge) && (fabs(cmax) < n.huge);
}

__device__ __forceinline__ void fastWarpReduce(FastAcc& r, int lane)
{
#pragma unroll
   for( int m = 16; m >= 1; m >>= 1 )
   {
      FastAcc o;
      o.minhi = __shfl_xor_sync(0xffffffffu, r.minhi, m);
      o.minlo = __shfl_xor_sync(0xffffffffu, r.minlo, m);
      o.maxhi = __shfl_xor_sync(0xffffffffu, r.maxhi, m);
      o.maxlo = __shfl_xor_sync(0xffffffffu, r.maxlo, m);
      o.maxdelta = __shfl_xor_sync(0xffffffffu, r.maxdelta, m);
      const bool upper = (lane & m) != 0;
      FastAcc x = upper ? o : r;
      const FastAcc y = upper ? r : o;
      dd_add_dd(x.minhi, x.minlo, y.minhi, y.minlo);
      dd_add_dd(x.maxhi, x.maxlo, y.maxhi, y.maxlo);
      x.maxdelta = fmax(x.maxdelta, y.maxdelta);
      r = x;
   }
}

// ---- exact, division-free early exit for the common row: all contributions finite, the row can neither tighten a
// ---- bound (tightenBounds gates :7057, :7081) nor be infeasible (propagateCons :7728).  Whatever this test cannot
// ---- decide goes to rowTighten, which restates the reference's rules in full.
__device__ __forceinline__ bool rowIsQuiet(const Num& n, const FastAcc& a, double lhs, double rhs)
{
   const double minact = a.minhi + a.minlo;
   const double maxact = a.maxhi + a.maxlo;
   // FeasGT(minact,rhs) needs (minact-rhs)/max(1,|minact|,|rhs|) > feastol, impossible if minact-rhs <= feastol/2
   if( minact - rhs > 0.5 * n.feastol || lhs - maxact > 0.5 * n.feastol )
      return false;
   if( fabs(a.maxdelta) <= n.feastol )
      return true;
   const double slack = isInf(n, rhs) ? n.inf : rhs - minact;
   const double surplus = isInf(n, -lhs) ? n.inf : maxact - lhs;
   return isLE(n, a.maxdelta, fmin(slack, surplus));
}

// gates, candidate pass and verdict of one row whose activities are known (elements first, first+step, ... of the
// calling thread); everything by value: the caller's state stays in registers
__device__ __noinline__ void rowTighten(const DevProblem* dp, double minhi, double minlo, double maxhi, double maxlo,
   double maxdelta, unsigned cnt, double lhs, double rhs, long long base, int stride, int first, int step, int len)
{
   const DevProblem& p = *dp;
   RowInfo ri;
   ri.acc.minhi = minhi;
   ri.acc.minlo = minlo;
   ri.acc.maxhi = maxhi;
   ri.acc.maxlo = maxlo;
   ri.acc.maxdelta = maxdelta;
   ri.acc.cnt = cnt;
   ri.lhs = lhs;
   ri.rhs = rhs;
   bool cutoff = false;
   if( rowGates(p.num, ri, len, cutoff) )
      rowCandidates(p, ri, base, stride, first, step, len, cutoff);
   if( cutoff || rowInfeasible(p.num, ri.acc, lhs, rhs) )
      p.ctrl->cutoff = 1;
}

// a row with infinite bounds / huge products, thread-per-row: full classification, sequentially
__device__ __noinline__ void rowSlowThread(const DevProblem* dp, double lhs, double rhs, long long base, int len)
{
   const DevProblem& p = *dp;
   RowAcc acc;
   accInit(acc);
   for( int k = 0; k < len; ++k )
   {
      const double a = p.vals[base + 32LL * k];
      const double2 b = p.bnd[p.cols[base + 32LL * k] & 0x7fffffff];
      accElem(p.num, acc, a, b.x, b.y);
   }
   rowTighten(dp, acc.minhi, acc.minlo, acc.maxhi, acc.maxlo, acc.maxdelta, acc.cnt, lhs, rhs, base, 32, 0, 1, len);
}

// the same for a warp-per-row row (called by all 32 lanes)
__device__ __noinline__ void rowSlowWarp(const DevProblem* dp, double lhs, double rhs, long long beg, int len, int lane)
{
   const DevProblem& p = *dp;
   RowAcc acc;
   accInit(acc);
   for( int idx = lane; idx < len; idx += 32 )
   {
      const double a = p.vals[beg + idx];
      const double2 b = p.bnd[p.cols[beg + idx] & 0x7fffffff];
      accElem(p.num, acc, a, b.x, b.y);
   }
   accWarpReduce(acc, lane);
   rowTighten(dp, acc.minhi, acc.minlo, acc.maxhi, acc.maxlo, acc.maxdelta, acc.cnt, lhs, rhs, beg, 1, lane, 32, len);
}

__device__ __forceinline__ void addRoundNnz(const DevProblem& p, unsigned long long nnzdone, int slot)
{
   if( nnzdone != 0 )
      atomicAdd(&p.ctrl->round_nnz[slot & (NNZ_SLOTS - 1)], nnzdone);
}

// ---- thread-per-row on SELL-32 slices: element k of the row of lane t sits at slice_off + 32 k + t -------------
// ---- persistent warps: warp w sweeps slices w, w + W, w + 2W, ...; the coefficients of chunk c+1 are in flight
// ---- while the bounds of chunk c are gathered
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
   if( act )
   {
      p.dirty[row] = 0;
      len = p.rowlen[row];
   }
   const long long base = p.sell_off[slice] + lane;
   const int maxlen = __reduce_max_sync(0xffffffffu, len);

   FastAcc acc;
   fastInit(acc);
   bool fast = true;
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
               fast &= fastElem(n, acc, a[k], b[k].x, b[k].y);
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
               fast &= fastElem(n, acc, a[k], b[k].x, b[k].y);
         }
      }
   }
   if( act )
   {
      const double2 sd = p.sides[row];
      if( !fast )
         rowSlowThread(p.self, sd.x, sd.y, base, len);
      else if( !rowIsQuiet(n, acc, sd.x, sd.y) )
         rowTighten(p.self, acc.minhi, acc.minlo, acc.maxhi, acc.maxlo, acc.maxdelta, 0u, sd.x, sd.y, base, 32, 0, 1, len);
      nnzdone += (unsigned)len;
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
         if( s0 + i * nw < nslices && row < p.nshort && p.dirty[row] != 0 )
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

// ---- thread-per-row on SELL-32 slices with a thread-private asynchronous pipeline (cp.async into shared memory) ----
// Latency is hidden by depth, not by occupancy: while chunk t is accumulated, the bounds of chunks t+1 .. t+D2-D1 are
// being gathered (16-byte cp.async.cg each) and the coefficients / column indices of the chunks up to t+D2 are
// streaming in.  Every thread copies and reads back only its own slots, so the pipeline needs no barrier at all:
// cp.async.wait_group orders a thread's own copies.  One warp walks slices w, w+W, ... in batches of 16 whose row
// lengths / flags / slice offsets are fetched up front.
__device__ __forceinline__ unsigned smemAddr(const void* ptr)
{
   return (unsigned)__cvta_generic_to_shared(ptr);
}
__device__ __forceinline__ void cpAsync4(unsigned dst, const void* src)
{
   asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cpAsync8(unsigned dst, const void* src)
{
   asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cpAsync16(unsigned dst, const void* src)
{
   asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cpAsyncCommit()
{
   asm volatile("cp.async.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void cpAsyncWait()
{
   asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

constexpr int ASYNC_BATCH = 16;   // slices per metadata batch

// position in the flat chunk sequence of a batch: slice i of the batch, first element c of the chunk
struct ChunkCursor
{
   int i;
   int c;
   int nslice;   // present slices passed so far (ring index of the row sides)
};

template <int CH, int D1, int D2, int THREADS>
__global__ void __launch_bounds__(THREADS) sweep_short_async_kernel(const DevProblem p)
{
   constexpr int NST = D2 + 1;
   extern __shared__ __align__(16) unsigned char smem_raw[];
   // [NST][CH][THREADS] of: double2 bounds | double coefficient | int column ; then [NST][THREADS] double2 sides
   double2* s_b = reinterpret_cast<double2*>(smem_raw);
   double* s_a = reinterpret_cast<double*>(s_b + NST * CH * THREADS);
   int* s_cj = reinterpret_cast<int*>(s_a + NST * CH * THREADS);
   double2* s_sd = reinterpret_cast<double2*>(s_cj + NST * CH * THREADS);

   const Num& n = p.num;
   const int tid = threadIdx.x;
   const int lane = tid & 31;
   const int gw = (blockIdx.x * THREADS + tid) >> 5;
   const int nw = (gridDim.x * THREADS) >> 5;
   const int nslices = (p.nshort + 31) >> 5;
   unsigned nnzdone = 0;

   for( int s0 = gw; s0 < nslices; s0 += ASYNC_BATCH * nw )
   {
      // ---- batch metadata: per lane the lengths of its rows (8 bits each), warp-uniform chunk extents, slice offsets
      unsigned lenpk[ASYNC_BATCH / 4];
      unsigned mlpk[ASYNC_BATCH / 4];
      unsigned actmask = 0u;      // per lane: my row of slice i is marked
      unsigned present = 0u;      // warp-uniform: slice i has a marked row
      long long soff = 0;         // lane i: element offset of slice i
#pragma unroll
      for( int q = 0; q < ASYNC_BATCH / 4; ++q )
      {
         lenpk[q] = 0u;
         mlpk[q] = 0u;
      }
      {
         int len[ASYNC_BATCH];
#pragma unroll
         for( int i = 0; i < ASYNC_BATCH; ++i )
         {
            const int slice = s0 + i * nw;
            const int row = slice * 32 + lane;
            len[i] = -1;
            if( slice < nslices && row < p.nshort && p.dirty[row] != 0 )
               len[i] = p.rowlen[row];
         }
         if( lane < ASYNC_BATCH && s0 + lane * nw < nslices )
            soff = p.sell_off[s0 + lane * nw];
#pragma unroll
         for( int i = 0; i < ASYNC_BATCH; ++i )
         {
            const bool act = len[i] >= 0;
            if( __any_sync(0xffffffffu, act) )
            {
               present |= 1u << i;
               if( act )
               {
                  actmask |= 1u << i;
                  p.dirty[(s0 + i * nw) * 32 + lane] = 0;
                  lenpk[i >> 2] |= (unsigned)len[i] << (8 * (i & 3));
                  nnzdone += (unsigned)len[i];
               }
               const int ml = __reduce_max_sync(0xffffffffu, act ? len[i] : 0);
               mlpk[i >> 2] |= (unsigned)(ml > 0 ? ml : 1) << (8 * (i & 3));
            }
         }
      }
      if( present == 0u )
         continue;

      // total chunks of the batch
      int total = 0;
#pragma unroll
      for( int i = 0; i < ASYNC_BATCH; ++i )
      {
         const int ml = (int)((mlpk[i >> 2] >> (8 * (i & 3))) & 0xffu);
         total += (ml + CH - 1) / CH;
      }

      auto lenOf = [&](int i) -> int {
         unsigned w = lenpk[0];
#pragma unroll
         for( int q = 1; q < ASYNC_BATCH / 4; ++q )
            w = ((i >> 2) == q) ? lenpk[q] : w;
         return (int)((w >> (8 * (i & 3))) & 0xffu);
      };
      auto mlOf = [&](int i) -> int {
         unsigned w = mlpk[0];
#pragma unroll
         for( int q = 1; q < ASYNC_BATCH / 4; ++q )
            w = ((i >> 2) == q) ? mlpk[q] : w;
         return (int)((w >> (8 * (i & 3))) & 0xffu);
      };
      auto advance = [&](ChunkCursor& cur) {
         cur.c += CH;
         if( cur.c >= mlOf(cur.i) )
         {
            cur.c = 0;
            ++cur.nslice;
            // next present slice (warp-uniform)
            const unsigned rest = present & ~((2u << cur.i) - 1u);
            cur.i = rest != 0u ? (__ffs(rest) - 1) : ASYNC_BATCH;
         }
      };

      ChunkCursor c1, c2, c3;
      c1.i = __ffs(present) - 1;
      c1.c = 0;
      c1.nslice = 0;
      c2 = c1;
      c3 = c1;

      FastAcc acc;
      fastInit(acc);
      bool fast = true;

      for( int t = 0; t < total + D2; ++t )
      {
         // ---- S1: coefficients + column indices of chunk t (and the sides of its row at the first chunk)
         if( t < total )
         {
            const int i = c1.i;
            const bool act = ((actmask >> i) & 1u) != 0u;
            const int len = act ? lenOf(i) : 0;
            const long long base = __shfl_sync(0xffffffffu, soff, i) + lane;
            const int st = t % NST;
#pragma unroll
            for( int k = 0; k < CH; ++k )
            {
               if( c1.c + k < len )
               {
                  cpAsync8(smemAddr(&s_a[(st * CH + k) * THREADS + tid]), p.vals + base + 32LL * (c1.c + k));
                  cpAsync4(smemAddr(&s_cj[(st * CH + k) * THREADS + tid]), p.cols + base + 32LL * (c1.c + k));
               }
            }
            if( c1.c == 0 && act )
               cpAsync16(smemAddr(&s_sd[(c1.nslice % NST) * THREADS + tid]), p.sides + ((s0 + i * nw) * 32 + lane));
            advance(c1);
         }
         cpAsyncCommit();

         // ---- S2: gather the bounds of chunk t - D1
         if( t >= D1 && t - D1 < total )
         {
            cpAsyncWait<2 * D1>();
            const int i = c2.i;
            const bool act = ((actmask >> i) & 1u) != 0u;
            const int len = act ? lenOf(i) : 0;
            const int st = (t - D1) % NST;
#pragma unroll
            for( int k = 0; k < CH; ++k )
            {
               if( c2.c + k < len )
               {
                  const int cj = s_cj[(st * CH + k) * THREADS + tid];
                  cpAsync16(smemAddr(&s_b[(st * CH + k) * THREADS + tid]), p.bnd + (cj & 0x7fffffff));
               }
            }
            advance(c2);
         }
         cpAsyncCommit();

         // ---- S3: accumulate chunk t - D2; at the last chunk of a slice finish its rows
         if( t >= D2 )
         {
            cpAsyncWait<2 * (D2 - D1)>();
            const int i = c3.i;
            const bool act = ((actmask >> i) & 1u) != 0u;
            const int len = act ? lenOf(i) : 0;
            const int st = (t - D2) % NST;
#pragma unroll
            for( int k = 0; k < CH; ++k )
            {
               if( c3.c + k < len )
               {
                  const double a = s_a[(st * CH + k) * THREADS + tid];
                  const double2 b = s_b[(st * CH + k) * THREADS + tid];
                  fast &= fastElem(n, acc, a, b.x, b.y);
               }
            }
            if( c3.c + CH >= mlOf(i) )
            {
               const long long base = __shfl_sync(0xffffffffu, soff, i) + lane;
               if( act )
               {
                  const double2 sd = s_sd[(c3.nslice % NST) * THREADS + tid];
                  if( !fast )
                     rowSlowThread(p.self, sd.x, sd.y, base, len);
                  else if( !rowIsQuiet(n, acc, sd.x, sd.y) )
                     rowTighten(p.self, acc.minhi, acc.minlo, acc.maxhi, acc.maxlo, acc.maxdelta, 0u, sd.x, sd.y, base, 32,
                        0, 1, len);
               }
               fastInit(acc);
               fast = true;
            }
            advance(c3);
         }
      }
      cpAsyncWait<0>();
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
   if( lane == 0 )
      p.dirty[row] = 0;
   const int len = p.rowlen[row];
   const long long beg = p.rowbeg[row];
   const double2 sd = p.sides[row];

   FastAcc acc;
   fastInit(acc);
   bool fast = true;
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
            fast &= fastElem(n, acc, a[k], b[k].x, b[k].y);
      }
   }
   if( __any_sync(0xffffffffu, !fast) )
      rowSlowWarp(p.self, sd.x, sd.y, beg, len, lane);
   else
   {
      fastWarpReduce(acc, lane);
      if( !rowIsQuiet(n, acc, sd.x, sd.y) )   // warp-uniform: every lane holds the same row state
         rowTighten(p.self, acc.minhi, acc.minlo, acc.maxhi, acc.maxlo, acc.maxdelta, 0u, sd.x, sd.y, beg, 1, lane, 32, len);
   }
   if( lane == 0 )
      nnzdone += (unsigned long long)len;
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
      const bool f = (mine < nrows) && p.dirty[row0 + mine] != 0;
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
// full classification in the loop (these rows are few); the reduction over the warps goes through shared memory
__global__ void __launch_bounds__(LONG_THREADS) sweep_long_kernel(const DevProblem p)
{
   __shared__ RowAcc s_acc[LONG_THREADS / 32];

   const int lane = threadIdx.x & 31;
   const int warp = threadIdx.x >> 5;
   const Num& n = p.num;
   const int row0 = p.nshort + p.nmedium;
   const int nrows = p.nrows - row0;

   for( int r = blockIdx.x; r < nrows; r += gridDim.x )
   {
      const int row = row0 + r;
      __syncthreads();                  // previous row done with s_acc and its dirty flag
      if( !p.dirty[row] )               // block-uniform: nobody clears the flag before the barrier below
         continue;
      __syncthreads();
      if( threadIdx.x == 0 )
         p.dirty[row] = 0;
      const int len = p.rowlen[row];
      const long long beg = p.rowbeg[row];

      RowAcc acc;
      accInit(acc);
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
               accElem(n, acc, a[k], b[k].x, b[k].y);
         }
      }
      accWarpReduce(acc, lane);
      if( lane == 0 )
         s_acc[warp] = acc;
      __syncthreads();
      acc = s_acc[0];
#pragma unroll 1
      for( int w = 1; w < LONG_THREADS / 32; ++w )
         accMerge(acc, s_acc[w]);

      const double2 sd = p.sides[row];
      rowTighten(p.self, acc.minhi, acc.minlo, acc.maxhi, acc.maxlo, acc.maxdelta, acc.cnt, sd.x, sd.y, beg, 1,
         threadIdx.x, LONG_THREADS, len);
      if( threadIdx.x == 0 )
         addRoundNnz(p, (unsigned long long)len, row);
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
   const long long e = p.colbeg[j + 1];
   for( long long q = p.colbeg[j]; q < e; ++q )
      p.dirty[p.colrows[q]] = 1;
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

// DENSE = false: only columns whose flag was raised by the sweep of this round are looked at (single GPU);
// DENSE = true : every column compares its (all-reduced) candidate keys with its bounds (rows sharded over ranks:
//                a key may have been moved by another rank)
template <bool DENSE, bool GRAPH>
__global__ void __launch_bounds__(256) apply_kernel(const DevProblem p, cudaGraphConditionalHandle handle)
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
   // warp-uniform trip count so that the log slots of a warp can be claimed with one atomic
   const int stride = gridDim.x * blockDim.x;
   const int niter = (p.ncols + stride - 1) / stride;
   for( int it = 0; it < niter; ++it )
   {
      const int j = it * stride + blockIdx.x * blockDim.x + threadIdx.x;
      bool lbchg = false;
      bool ubchg = false;
      double2 nb = make_double2(0.0, 0.0);
      int nc = 0;
      if( j < p.ncols )
      {
         bool look;
         if( DENSE )
            look = true;
         else
            look = p.colflag[j] != 0;
         if( look )
         {
            p.colflag[j] = 0;
            nc = applyColumn(p, j, nb, lbchg, ubchg);
            if( nc > 0 )
               markColumnRows(p, j);
         }
      }
      mychg += nc;
      if( logcap > 0 )
      {
         const unsigned any = __ballot_sync(0xffffffffu, nc > 0);
         if( any != 0u )
         {
            // exclusive prefix of nc over the warp
            int incl = nc;
#pragma unroll
            for( int d = 1; d < 32; d <<= 1 )
            {
               const int t = __shfl_up_sync(0xffffffffu, incl, d);
               if( lane >= d )
                  incl += t;
            }
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            unsigned long long base = 0;
            if( lane == 0 )
               base = atomicAdd(&c->logcount, (unsigned long long)total);
            base = __shfl_sync(0xffffffffu, base, 0);
            unsigned long long pos = base + (unsigned long long)(incl - nc);
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
      p.dirty[r] = 1;
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
      p.dirty[r] = 1;
}

// start of a gpulin_propagate call: reset the loop state
__global__ void begin_kernel(Ctrl* c)
{
   c->round = 0;
   c->cont = 1;
   c->status = 0;
   c->cutoff = 0;
   c->ticket = 0;
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
