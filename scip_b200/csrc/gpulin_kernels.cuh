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
constexpr int MAX_CLASSES = 12;
constexpr int MAX_HIST = 1024;       // rounds with recorded per-round statistics
constexpr int SHORT_MAXLEN = 32;     // thread-per-row up to this length
constexpr int MEDIUM_MAXLEN = 1024;  // warp-per-row up to this length

// row classes of the fused sweep kernel
enum ClassKind { CK_T4 = 0, CK_T8, CK_T16, CK_T32, CK_W2, CK_W4, CK_W8, CK_W16, CK_W32 };

struct ClassTable
{
   int n;
   int kind[MAX_CLASSES];
   int row0[MAX_CLASSES];     // first row (permuted numbering) of the class
   int nrows[MAX_CLASSES];
   int block0[MAX_CLASSES + 1];
};

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
   unsigned long long round_nnz;    // nonzeros swept in the running round
   unsigned long long total_nchg;
   unsigned long long total_nnz;
   unsigned long long t_start;      // %globaltimer at the start of the call
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
   // rows in permuted order: short rows (SELL-32 slices, ascending length), then medium, then long (CSR)
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
   Num                 num;
};

__device__ __forceinline__ unsigned long long globaltimer()
{
   unsigned long long t;
   asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
   return t;
}

// streaming loads of the matrix: read once per round, keep them out of L1
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

// ---- candidate pass for one nonzero that passed the slack test (or of a row on the general path): re-read it
__device__ __forceinline__ void candidateAt(const DevProblem& p, const RowInfo& ri, long long pos, bool& cutoff)
{
   const double a = p.vals[pos];
   const int cj = p.cols[pos];
   const int j = cj & 0x7fffffff;
   const double2 b = p.bnd[j];
   Sink s;
   s.cand = p.cand;
   s.colflag = p.colflag;
   candidates(p.num, s, ri, a, j, cj < 0, b.x, b.y, cutoff);
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

// ---- thread-per-row on a SELL-32 slice: element k of the row of lane t sits at slice_off + 32 k + t ------------
template <int MAXLEN>
__device__ __forceinline__ void sweepThreadRow(const DevProblem& p, int row, int& nnzdone)
{
   if( !p.dirty[row] )
      return;
   p.dirty[row] = 0;
   const int len = p.rowlen[row];
   const long long base = p.sell_off[row >> 5] + (row & 31);
   const Num& n = p.num;
   nnzdone += len;

   double alpha[MAXLEN];
   RowInfo ri;
   accInit(ri.acc);

   constexpr int CH = MAXLEN < 8 ? MAXLEN : 8;
#pragma unroll
   for( int c = 0; c < MAXLEN; c += CH )
   {
      if( c < len )
      {
         double a[CH];
         int cj[CH];
         double2 b[CH];
#pragma unroll
         for( int k = 0; k < CH; ++k )
         {
            if( c + k < len )
            {
               a[k] = ldStream(p.vals + base + 32LL * (c + k));
               cj[k] = ldStream(p.cols + base + 32LL * (c + k));
            }
         }
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
            {
               accElem(n, ri.acc, a[k], b[k].x, b[k].y);
               alpha[c + k] = fabs(a[k]) * (b[k].y - b[k].x);
            }
         }
      }
   }

   const double2 sd = p.sides[row];
   ri.lhs = sd.x;
   ri.rhs = sd.y;
   bool cutoff = false;
   if( rowGates(n, ri, len, cutoff) )
   {
      if( ri.easy )
      {
         const double thr = slackThreshold(n, ri.force);
#pragma unroll
         for( int k = 0; k < MAXLEN; ++k )
         {
            if( k < len && passesSlackTest(ri, alpha[k], thr) )
               candidateAt(p, ri, base + 32LL * k, cutoff);
         }
      }
      else
      {
         for( int k = 0; k < len; ++k )
            candidateAt(p, ri, base + 32LL * k, cutoff);
      }
   }
   if( cutoff || rowInfeasible(n, ri.acc, ri.lhs, ri.rhs) )
      p.ctrl->cutoff = 1;
}

// ---- warp-per-row on CSR: lane t holds elements t, t+32, ... (at most K per lane) ------------------------------
template <int K>
__device__ __forceinline__ void sweepWarpRow(const DevProblem& p, int row, int lane, int& nnzdone)
{
   int isdirty = (lane == 0) ? (int)p.dirty[row] : 0;
   isdirty = __shfl_sync(0xffffffffu, isdirty, 0);
   if( !isdirty )
      return;
   if( lane == 0 )
      p.dirty[row] = 0;
   const int len = p.rowlen[row];
   const long long beg = p.rowbeg[row];
   const Num& n = p.num;
   if( lane == 0 )
      nnzdone += len;

   double alpha[K];
   RowInfo ri;
   accInit(ri.acc);

   constexpr int CH = K < 4 ? K : 4;
#pragma unroll
   for( int c = 0; c < K; c += CH )
   {
      if( c * 32 < len )
      {
         double a[CH];
         int cj[CH];
         double2 b[CH];
#pragma unroll
         for( int k = 0; k < CH; ++k )
         {
            const int idx = (c + k) * 32 + lane;
            if( idx < len )
            {
               a[k] = ldStream(p.vals + beg + idx);
               cj[k] = ldStream(p.cols + beg + idx);
            }
         }
#pragma unroll
         for( int k = 0; k < CH; ++k )
         {
            if( (c + k) * 32 + lane < len )
               b[k] = p.bnd[cj[k] & 0x7fffffff];
         }
#pragma unroll
         for( int k = 0; k < CH; ++k )
         {
            if( (c + k) * 32 + lane < len )
            {
               accElem(n, ri.acc, a[k], b[k].x, b[k].y);
               alpha[c + k] = fabs(a[k]) * (b[k].y - b[k].x);
            }
         }
      }
   }
   accWarpReduce(ri.acc, lane);

   const double2 sd = p.sides[row];
   ri.lhs = sd.x;
   ri.rhs = sd.y;
   bool cutoff = false;
   if( rowGates(n, ri, len, cutoff) )   // warp-uniform: every lane holds the same row state
   {
      if( ri.easy )
      {
         const double thr = slackThreshold(n, ri.force);
#pragma unroll
         for( int k = 0; k < K; ++k )
         {
            const int idx = k * 32 + lane;
            if( idx < len && passesSlackTest(ri, alpha[k], thr) )
               candidateAt(p, ri, beg + idx, cutoff);
         }
      }
      else
      {
         for( int idx = lane; idx < len; idx += 32 )
            candidateAt(p, ri, beg + idx, cutoff);
      }
   }
   if( cutoff || (lane == 0 && rowInfeasible(n, ri.acc, ri.lhs, ri.rhs)) )
      p.ctrl->cutoff = 1;
}

// ---- fused sweep over the short and medium classes ------------------------------------------------------------
__global__ void __launch_bounds__(SWEEP_THREADS) sweep_rows_kernel(const DevProblem p, const ClassTable ct)
{
   __shared__ int s_nnz;
   if( threadIdx.x == 0 )
      s_nnz = 0;
   __syncthreads();

   int c = 0;
   while( c + 1 < ct.n && (int)blockIdx.x >= ct.block0[c + 1] )
      ++c;
   const int lb = blockIdx.x - ct.block0[c];
   const int kind = ct.kind[c];
   const int lane = threadIdx.x & 31;
   int nnzdone = 0;

   if( kind <= CK_T32 )
   {
      const int local = lb * SWEEP_THREADS + threadIdx.x;
      if( local < ct.nrows[c] )
      {
         const int row = ct.row0[c] + local;
         switch( kind )
         {
         case CK_T4: sweepThreadRow<4>(p, row, nnzdone); break;
         case CK_T8: sweepThreadRow<8>(p, row, nnzdone); break;
         case CK_T16: sweepThreadRow<16>(p, row, nnzdone); break;
         default: sweepThreadRow<32>(p, row, nnzdone); break;
         }
      }
   }
   else
   {
      const int local = lb * (SWEEP_THREADS / 32) + (threadIdx.x >> 5);
      if( local < ct.nrows[c] )   // warp-uniform
      {
         const int row = ct.row0[c] + local;
         switch( kind )
         {
         case CK_W2: sweepWarpRow<2>(p, row, lane, nnzdone); break;
         case CK_W4: sweepWarpRow<4>(p, row, lane, nnzdone); break;
         case CK_W8: sweepWarpRow<8>(p, row, lane, nnzdone); break;
         case CK_W16: sweepWarpRow<16>(p, row, lane, nnzdone); break;
         default: sweepWarpRow<32>(p, row, lane, nnzdone); break;
         }
      }
   }
   __syncwarp();
   nnzdone = __reduce_add_sync(0xffffffffu, nnzdone);
   if( lane == 0 && nnzdone != 0 )
      atomicAdd(&s_nnz, nnzdone);
   __syncthreads();
   if( threadIdx.x == 0 && s_nnz != 0 )
      atomicAdd(&p.ctrl->round_nnz, (unsigned long long)s_nnz);
}

// ---- block-per-row for long rows: alpha staged in shared memory when it fits ---------------------------------------
__global__ void __launch_bounds__(LONG_THREADS) sweep_long_kernel(const DevProblem p, int row0, int nrows, int smemcap)
{
   extern __shared__ double s_alpha[];
   __shared__ RowAcc s_acc[LONG_THREADS / 32];

   const int lane = threadIdx.x & 31;
   const int warp = threadIdx.x >> 5;
   const Num& n = p.num;

   for( int r = blockIdx.x; r < nrows; r += gridDim.x )
   {
      const int row = row0 + r;
      __syncthreads();                  // previous row done with s_acc / s_alpha / dirty flag
      if( !p.dirty[row] )      // block-uniform: nobody clears the flag before the barrier below
         continue;
      __syncthreads();
      if( threadIdx.x == 0 )
         p.dirty[row] = 0;
      const int len = p.rowlen[row];
      const long long beg = p.rowbeg[row];
      const bool staged = len <= smemcap;

      RowInfo ri;
      accInit(ri.acc);
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
            const int idx = i0 + k * LONG_THREADS + threadIdx.x;
            if( idx < len )
            {
               accElem(n, ri.acc, a[k], b[k].x, b[k].y);
               if( staged )
                  s_alpha[idx] = fabs(a[k]) * (b[k].y - b[k].x);
            }
         }
      }
      accWarpReduce(ri.acc, lane);
      if( lane == 0 )
         s_acc[warp] = ri.acc;
      __syncthreads();
      ri.acc = s_acc[0];
#pragma unroll 1
      for( int w = 1; w < LONG_THREADS / 32; ++w )
         accMerge(ri.acc, s_acc[w]);

      const double2 sd = p.sides[row];
      ri.lhs = sd.x;
      ri.rhs = sd.y;
      bool cutoff = false;
      if( rowGates(n, ri, len, cutoff) )   // block-uniform
      {
         if( ri.easy && staged )
         {
            const double thr = slackThreshold(n, ri.force);
            for( int idx = threadIdx.x; idx < len; idx += LONG_THREADS )
            {
               if( passesSlackTest(ri, s_alpha[idx], thr) )
                  candidateAt(p, ri, beg + idx, cutoff);
            }
         }
         else
         {
            for( int idx = threadIdx.x; idx < len; idx += LONG_THREADS )
               candidateAt(p, ri, beg + idx, cutoff);
         }
      }
      if( cutoff || (threadIdx.x == 0 && rowInfeasible(n, ri.acc, ri.lhs, ri.rhs)) )
         p.ctrl->cutoff = 1;
      if( threadIdx.x == 0 )
         atomicAdd(&p.ctrl->round_nnz, (unsigned long long)len);
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
   if( r < MAX_HIST )
   {
      c->hist_time[r] = globaltimer();
      c->hist_nnz[r] = c->round_nnz;
      c->hist_nchg[r] = nchg;
   }
   c->total_nchg += nchg;
   c->total_nnz += c->round_nnz;
   c->round_nchg = 0;
   c->round_nnz = 0;
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
   c->round_nnz = 0;
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
