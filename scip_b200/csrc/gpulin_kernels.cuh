// gpulin_kernels.cuh -- the propagation round of the linear bound propagation path as sm_100a kernels.
//
// One round (= one sweep of the reference's consPropLinear over its marked rows, cons_linear.c:16126-16195, run as a
// synchronous Jacobi step) is three launches:
//
//   sweep_stream_kernel (+ sweep_long_kernel for rows > STREAM_MAXLEN)      THE HOT KERNEL: an interval filter.
//        All rows of 1..STREAM_MAXLEN nonzeros lie concatenated in one CSR stream that is cut into tiles of 256
//        nonzeros.  A warp owns a contiguous range of tiles; per tile every lane reads 8 consecutive coefficients and
//        column indices with 16-byte vector loads, gathers the 8 bound pairs (8 independent 16-byte loads in flight per
//        lane), and the per-row sums minact/maxact/maxdelta/cabs come out of a segmented scan over the tile (row ends
//        are a precomputed bit per nonzero; rows may span tiles and warp ranges).  9 fp64 instructions per nonzero,
//        the matrix is streamed once.  A row that is CLEARLY QUIET (see rowClearlyQuiet) is finished; every other
//        marked row goes to a work list.
//   exact_rows_kernel      the reference's rules in full for the rows on the work lists: double-double activities with
//        inf/huge counters, gates of tightenBounds, candidate bounds per nonzero (tightenVarBoundsEasy /
//        tightenVarBounds), commit filter, atomicMin on the int64 keys, row verdict.
//   apply_kernel           for every column whose key moved: accept the new bounds, detect crossing bounds, log the
//        change, mark the rows of the column (CSC) for the next round -- the counterpart of eventExecLinear's
//        SCIPmarkConsPropagate (:17229); the last block to finish runs the loop control (propagateDomains,
//        solve.c:766) and sets the CUDA-graph WHILE condition.
#pragma once

#include "gpulin_device.cuh"
#include "gpulin_ranged.cuh"

namespace gpl {

constexpr int SWEEP_THREADS = 128;
constexpr int LONG_THREADS = 512;
constexpr int MAX_HIST = 1024;        // rounds with recorded per-round statistics
constexpr int MAX_TRACE = 256;        // kernel starts with a recorded time stamp per call
enum TraceId { TR_BEGIN = 1, TR_SWEEP_SELL, TR_SWEEP_STREAM, TR_SWEEP_LONG, TR_EXACT, TR_PUSH, TR_PUSH_END, TR_MERGE,
               TR_MERGE_READY, TR_APPLY, TR_SPARSE, TR_SPARSE_END };
constexpr int TILE = 256;             // nonzeros per tile of the CSR stream = 32 lanes x TPL
constexpr int TPL = 8;                // nonzeros per lane and tile
constexpr int STREAM_MAXLEN = 4096;   // longer rows (and empty rows) are swept block-per-row
constexpr int SHORT_MAXLEN = 32;      // exact kernel: rows up to this length are handled by groups of 8 lanes
constexpr int NNZ_SLOTS = 64;         // spread counters of the nonzeros swept in a round
constexpr int MARKCAP = 1 << 15;      // capacity of a mark list; rounds with more marked rows take the dense sweeps
constexpr int COL_INTEGRAL = (int)0x80000000u;   // cols[] word: the column is integral
constexpr int COL_NEGCOEF = 0x40000000;          // cols[] word: the coefficient of this nonzero is negative
constexpr int COL_MASK = 0x3fffffff;             // cols[] word: the column index

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
   unsigned int       nchgcols;     // columns on the change list of the running round
   unsigned int       epoch;        // exchanges with the peer ranks done so far (never reset; the same number on every rank)
   int                peererror;    // a peer rank did not deliver its candidates in time
   unsigned int       pushticket;   // peer_push_kernel: blocks finished
   unsigned long long logcount;     // entries produced
   unsigned long long round_nchg;   // accepted bound changes of the running round
   unsigned long long total_nchg;
   unsigned long long total_nnz;
   unsigned long long t_start;      // %globaltimer at the start of the call
   unsigned int       nexact[4];    // rows the filter sweeps handed to the exact kernel in the running round, per list
                                    // ([0..2]: xlist per bin, [3]: flist, the short rows whose exact activities came along)
   alignas(16) unsigned int nmark[2][4];  // (16-byte aligned: read as one word) rows marked (per bin) into mark list buffer 0 / 1; marking writes buffer mb^1
   unsigned int       mb;           // the buffer the last apply filled = what a sparse round reads
   unsigned int       resume;       // the next begin_kernel continues the call small_rounds started (round count, totals, log)
   unsigned int       nsparse;      // rounds done by sparse_rounds_kernel in this call (statistics)
   unsigned int       poisoned;     // probing worker: its state is not "node + change log" any more (see probe_kernel)
   unsigned int       nfastrows;    // rows that took the thread-per-row phase of the exact kernel in this call (statistics)
   unsigned long long round_nnz[NNZ_SLOTS];  // nonzeros swept in the running round (sum over the slots)
   unsigned long long hist_time[MAX_HIST];   // %globaltimer at the end of each round
   unsigned long long hist_nnz[MAX_HIST];
   unsigned long long hist_nchg[MAX_HIST];
   unsigned int       ntrace;                // kernel-start stamps of the call (the first MAX_TRACE): (id << 56) | %globaltimer
   unsigned long long trace[MAX_TRACE];
   unsigned long long hist_push[MAX_HIST];   // several GPUs: %globaltimer when the exchange of the round started (0: none) ...
   unsigned long long hist_wait[MAX_HIST];   // ... and how long the merge waited for the slowest peer [ns]
};

struct ChangeRec      // == gpulin_change
{
   int    var;
   int    round;
   double newbound;
   int    is_upper;
   int    reserved;
};

// exact activities that come along with a row the filter sweep hands over (flist / facc).  If every coefficient of a row is
// an integer, every column is of integral type and every bound of those columns is an integer (ROWLEN_INT, fracflag), all
// products a * bound are integers; as long as their absolute sum stays below 2^40 EVERY partial sum of the filter's plain
// fp64 accumulation (in whatever order, also in the midpoint / half-width form: multiples of 1/2) is exact -- and then the
// double-double accumulation of the exact rules yields the same value with a zero low word (dd_add: s = hi + c is exact, so
// e = 0).  Such a row skips the activity pass of the exact rules: gates and candidates start from these three numbers.
struct alignas(32) FastAcc
{
   double minact;
   double maxact;
   double maxdelta;
   int    len;        // the header of the row comes along (the sweep has it in registers): one round trip less in the
   int    off32;      // dependent chain of the phase that takes these rows (fastShortPhase)
   double lhs;
   double rhs;
   double spare[2];
};
constexpr int FAST_CABSHI = 0x42700000;        // high word of 2^40: the bound on the absolute sums
constexpr float FAST_MAXDELTA = 4194304.0f;    // 2^22: below it the fp32 maximum (of multiples of 1/2, rounded up) is exact

// state of a column as the unit slices of the bit-table sweep see it (colstate[], kept by noteBounds)
constexpr unsigned char CS_FREE = 0;     // (0,1)
constexpr unsigned char CS_FIX0 = 1;     // (0,0)
constexpr unsigned char CS_FIX1 = 2;     // (1,1)
constexpr unsigned char CS_OTHER = 3;    // anything else: the bound pair is gathered

struct DevProblem
{
   int                 nrows;
   int                 ncols;
   int                 nsell;      // rows [0,nsell): 1..32 nonzeros, SELL-32 slices sorted by length (thread-per-row sweep)
   int                 nsellunit;  // rows [0,nsellunit) (a multiple of 32): every coefficient is +1 or -1 -- the filter sweep
                                   // takes the sign from the column word and does not read their values at all
   int                 nsx;        // rows [nsell,nsx): 33..STREAM_MAXLEN nonzeros, the CSR stream in the caller's order;
                                   // rows [nsx,nrows): longer or empty, swept block-per-row
   int                 ntiles;     // tiles of the stream (its storage is padded with zero coefficients to a whole tile)
   long long           streambase; // element offset of the first tile of the stream (a multiple of TILE)
   const long long*    sell_off;   // per SELL slice: element offset; element k of row r sits at sell_off[r>>5] + (r&31) + 32 k
   const int*          sell_off32; // sell_off / 32 (the offsets are multiples of 32)
   // rows (permuted numbering)
   const long long*    rowbeg;     // element offset of the first nonzero
   const int*          rowlen;     // length | ROWLEN_EXACT | ROWLEN_INT
   const double2*      sides;      // (lhs, rhs)
   unsigned char*      dirty;      // marked for propagation
   const double*       vals;
   const int*          cols;       // column index | COL_NEGCOEF (coefficient < 0) | COL_INTEGRAL
   // tiles
   const int*          tile_row0;  // first row that ends at or behind the first nonzero of the tile
   const unsigned char* endmask;   // per tile and lane: bit i = nonzero 8*lane+i is the last of its row
   unsigned char*      tileflag;   // a marked row starts in this tile
   int*                xlist;      // rows handed to exact_rows_kernel, one list per bin: [0,nsell) [nsell,nsx) [nsx,nrows)
   int*                flist;      // rows of the thread-per-row class handed over TOGETHER WITH their exact activities (FastAcc) ...
   FastAcc*            facc;       // ... which sit at the same position of this array
   unsigned*           fracflag;   // != 0: a column of integral type has (had) a bound that is not an integer -- no FastAcc
   const unsigned char* coltype;   // per column: != 0 integral type
   int*                marklist;   // rows marked by the apply step: [buffer 0|1][bin 0..2][MARKCAP]; complete unless a
                                   // count exceeds MARKCAP (the dirty flags are the ground truth, the lists a shortcut)
   // columns
   const double2*      bnd;        // (lb, ub) at round start
   double2*            bndf;       // ((lb+ub)/2, (ub-lb)/2) per column: what sweep_sell_bits_kernel gathers instead of bnd
   unsigned char*      colstate;   // CS_* per column
   unsigned*           freebits;   // bit per column: bnd == (0,1) exactly; nfreewords words (a multiple of 4) are staged into
   int                 nfreewords; // shared memory by sweep_sell_bits_kernel, they cover the columns [0, nfreecols)
   int                 nfreecols;
   long long*          cand;       // 2*ncols (+2) candidate keys, see Sink
   unsigned*           colbits;    // bit per column: on the change list
   int*                chglist;    // columns a candidate reached in the running round
   // column -> rows (permuted row ids)
   const long long*    colbeg;
   const int*          colrows;
   Ctrl*               ctrl;
   ChangeRec*          log;
   const PeerTable*    peers;      // NULL: single GPU
   // the share of this rank in a dense round (single GPU: everything): every nranks-th SELL slice of the unit class and
   // of the others (rank, rank + nranks, ...: neighbouring slices hold rows of the same length and the same kind, so the
   // ranks get equal work in the filter AND in the exact rules), tiles [st0,st1) of the stream, every nranks-th
   // block-per-row row
   int                 st0, st1;
   int                 nranks, rank;
   unsigned            markall_min; // an apply step with at least this many changed columns marks ALL rows (see apply_kernel)
   unsigned            fastmin;     // the rows of flist are finished by fast_rows_kernel if there are more than this ...
   int                 fastround;   // ... in a round that launches it: the first round of a call (the one round with work for it;
                                    // in the later rounds of the loop its empty launch would cost 3 - 5 us each)
   RangedRows          rr;          // ranged-row propagation (gpulin_set_rangedrow); rr.n == 0: off
   Num                 num;
};

__device__ __forceinline__ unsigned long long globaltimer()
{
   unsigned long long t;
   asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
   return t;
}
__device__ __forceinline__ void traceStamp(Ctrl* c, int id)
{
   const unsigned i = atomicAdd(&c->ntrace, 1u);
   if( i < (unsigned)MAX_TRACE )
      c->trace[i] = ((unsigned long long)id << 56) | (globaltimer() & 0x00ffffffffffffffull);
}
#define TRACE_KERNEL_START(c, id) do { if( blockIdx.x == 0 && threadIdx.x == 0 ) traceStamp((c), (id)); } while( 0 )

// time stamp of a kernel start / phase (called by one thread)
// streaming loads of the matrix: read once per round, evict-first, 16 bytes per request
__device__ __forceinline__ double2 ldStream2(const double* p)
{
   return __ldcs(reinterpret_cast<const double2*>(p));
}
__device__ __forceinline__ int4 ldStream4(const int* p)
{
   return __ldcs(reinterpret_cast<const int4*>(p));
}
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
   return force ? dmin(n.eps, n.sumeps) : n.sumeps;
}

// ---- candidate pass over the elements first, first+step, ... < len of a row that passed the gates of tightenBounds
// ---- (rare once the bounds have settled): the row is read a second time, from L2.  Easy rows only look at nonzeros
// ---- whose alpha = |a| (ub - lb) exceeds the slack (tightenVarBoundsEasy :5474/:5566).  Few nonzeros pass that test and
// ---- the candidate rules are long, so the passing ones are COMPACTED first: every warp collects their positions in a
// ---- small queue in shared memory and works through the queue 32 at a time, one nonzero per lane -- the rules run with
// ---- full warps, and there is one copy of them instead of one per unroll slot (the instruction cache and 1-2 active
// ---- lanes per pass were what made the exact kernel slow).  `first - lane` must be uniform over the warp.
constexpr int CAND_QCAP = 160;      // < 32 left over + 4 x 32 new positions
struct CandQueue
{
   int k[CAND_QCAP];
};

// ALPHA (warp per row: first == lane, step == 32): alpha[q] = |a| (ub - lb) of the elements lane + 32 q, q < 8, was kept
// from the activity pass -- the slack test of the first 256 nonzeros needs no load at all, only the few nonzeros that
// pass it are fetched again (from L1)
constexpr int ALPHA_SLOTS = 8;
template <bool ALPHA, bool LISTED>
__device__ __forceinline__ void rowCandidates(const DevProblem& p, const RowInfo& ri, long long base, int stride,
   int first, int step, int len, bool& cutoff, CandQueue& queue, const double (&alpha)[ALPHA_SLOTS])
{
   const Num& n = p.num;
   const double thr = slackThreshold(n, ri.force);
   const int lane = threadIdx.x & 31;
   const unsigned below = (1u << lane) - 1u;
   Sink s;
   s.cand = p.cand;
   s.colbits = p.colbits;
   s.chglist = p.chglist;
   s.nchgcols = &p.ctrl->nchgcols;
   s.listed = LISTED;
   int cnt = 0;                        // positions in the queue (warp-uniform)
   int kb = first - lane;              // warp-uniform
   for( ;; )
   {
      if( kb < len )
      {
         const int k0 = kb + lane;
         double al[4];
         if( ALPHA && kb < 32 * ALPHA_SLOTS )
         {
#pragma unroll
            for( int q = 0; q < 4; ++q )
               al[q] = kb == 0 ? alpha[q] : alpha[4 + q];
         }
         else
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
                  b[q] = p.bnd[cj[q] & COL_MASK];
            }
#pragma unroll
            for( int q = 0; q < 4; ++q )
               al[q] = k0 + q * step < len ? fabs(a[q]) * (b[q].y - b[q].x) : 0.0;
         }
#pragma unroll
         for( int q = 0; q < 4; ++q )
         {
            const bool pass = k0 + q * step < len && (!ri.easy || passesSlackTest(ri, al[q], thr));
            const unsigned m = __ballot_sync(0xffffffffu, pass);
            if( pass )
               queue.k[cnt + __popc(m & below)] = k0 + q * step;
            cnt += __popc(m);
         }
         kb += 4 * step;
      }
      const bool last = kb >= len;
      int head = 0;
      while( cnt - head >= 32 || (last && cnt > head) )
      {
         const int take = min(cnt - head, 32);
         __syncwarp();
         bool touched = false;
         int col = 0;
         if( lane < take )
         {
            const int k = queue.k[head + lane];
            const double a = p.vals[base + (long long)stride * k];
            const int cj = p.cols[base + (long long)stride * k];
            col = cj & COL_MASK;
            const double2 b = p.bnd[col];
            candidates(n, s, ri, a, col, cj < 0, b.x, b.y, cutoff, touched);
         }
         const bool firsttouch = touched && raiseColumnBit(s, col);
         listChangedColumn(s, col, firsttouch);
         head += take;
      }
      if( head > 0 )
      {
         // what is left (fewer than 32 positions) moves to the front
         const int rest = cnt - head;
         int mv = 0;
         if( lane < rest )
            mv = queue.k[head + lane];
         __syncwarp();
         if( lane < rest )
            queue.k[lane] = mv;
         cnt = rest;
      }
      if( last )
         break;
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
constexpr int ROWLEN_INT = 0x20000000;     // flag in rowlen[]: every coefficient is an integer (|a| <= 2^20) and every column
                                           // of the row is of integral type -- see FastAcc
constexpr int ROWLEN_MASK = ~(ROWLEN_EXACT | ROWLEN_INT);

struct LeanAcc
{
   double minact;     // plain fp64 sums
   double maxact;
   float  maxdelta;   // max_k (cmax_k - cmin_k), rounded UP to fp32 (exact for integers below 2^24)
   int    cabshi;     // high word of max_k max(|a_k lb_k|, |a_k ub_k|): an upper bound is (cabshi + 1) << 32
};

__device__ __forceinline__ void leanInit(LeanAcc& r)
{
   r.minact = r.maxact = 0.0;
   r.maxdelta = 0.0f;
   r.cabshi = 0;
}

// 17 instructions per nonzero: the sign of the coefficient comes from its high word, the two maxima are a float and
// an integer maximum (non-negative IEEE doubles are ordered like their bit patterns); fp64 min/max would cost ~8
// instructions each on this architecture
__device__ __forceinline__ void leanElem(LeanAcc& r, double a, double l, double u)
{
   const double al = a * l;
   const double au = a * u;
   const bool pos = __double2hiint(a) >= 0;
   const double cmin = pos ? al : au;      // = min(al, au) for lb <= ub
   const double cmax = pos ? au : al;
   r.minact += cmin;
   r.maxact += cmax;
   r.maxdelta = fmaxf(r.maxdelta, __double2float_ru(cmax - cmin));
   r.cabshi = max(r.cabshi, max(__double2hiint(al) & 0x7fffffff, __double2hiint(au) & 0x7fffffff));
}

__device__ __forceinline__ void leanAdd(LeanAcc& r, const LeanAcc& o)
{
   r.minact += o.minact;
   r.maxact += o.maxact;
   r.maxdelta = fmaxf(r.maxdelta, o.maxdelta);
   r.cabshi = max(r.cabshi, o.cabshi);
}

__device__ __forceinline__ LeanAcc leanShflUp(const LeanAcc& s, int d)
{
   LeanAcc o;
   o.minact = __shfl_up_sync(0xffffffffu, s.minact, d);
   o.maxact = __shfl_up_sync(0xffffffffu, s.maxact, d);
   o.maxdelta = __shfl_up_sync(0xffffffffu, s.maxdelta, d);
   o.cabshi = __shfl_up_sync(0xffffffffu, s.cabshi, d);
   return o;
}

__device__ __forceinline__ LeanAcc leanShfl(const LeanAcc& s, int src)
{
   LeanAcc o;
   o.minact = __shfl_sync(0xffffffffu, s.minact, src);
   o.maxact = __shfl_sync(0xffffffffu, s.maxact, src);
   o.maxdelta = __shfl_sync(0xffffffffu, s.maxdelta, src);
   o.cabshi = __shfl_sync(0xffffffffu, s.cabshi, src);
   return o;
}

__device__ __forceinline__ void leanWarpReduce(LeanAcc& r)
{
#pragma unroll
   for( int m = 16; m >= 1; m >>= 1 )
   {
      r.minact += __shfl_xor_sync(0xffffffffu, r.minact, m);
      r.maxact += __shfl_xor_sync(0xffffffffu, r.maxact, m);
      r.maxdelta = fmaxf(r.maxdelta, __shfl_xor_sync(0xffffffffu, r.maxdelta, m));
      r.cabshi = max(r.cabshi, __shfl_xor_sync(0xffffffffu, r.cabshi, m));
   }
}

__device__ __forceinline__ bool rowClearlyQuiet(const Num& n, const LeanAcc& r, int len, double lhs, double rhs)
{
   // upper bound of the largest |product|; an infinite bound (|b| >= 1e20, with |a| >= hugeval/infinity by the
   // ROWLEN_EXACT flag), a huge product or a NaN shows up here
   if( r.cabshi >= 0x7fe00000 )
      return false;
   const double cabs = __hiloint2double(r.cabshi + 1, 0);
   if( !(cabs < n.huge) )
      return false;
   const double dl = (double)len;      // (in double: len * len overflows an int from 46341 nonzeros on)
   const double E = (dl * dl + 8.0) * 2.3e-16 * cabs;
   // verdict: FeasGT(minact,rhs) needs (minact-rhs)/max(1,|minact|,|rhs|) > feastol, impossible if minact-rhs <= feastol/4
   if( (r.minact - rhs) + E > 0.25 * n.feastol || (lhs - r.maxact) + E > 0.25 * n.feastol )
      return false;
   const double md = (double)r.maxdelta + E;
   if( md <= 0.5 * n.feastol )
      return true;                // all variables fixed (:7057)
   const double slack = isInf(n, rhs) ? n.inf : rhs - r.minact;
   const double surplus = isInf(n, -lhs) ? n.inf : r.maxact - lhs;
   const double m = dmin(slack, surplus);
   // the gate maxdelta <= min(slack, surplus) + eps (:7081) with every rounding in its disfavour
   return md + 2.0 * E + 4.5e-16 * fabs(m) - m <= 0.5 * n.eps;
}

// gates, candidate pass and verdict of one row whose exact activities are known (elements first, first+step, ... of
// the calling thread)
template <bool ALPHA, bool LISTED>
__device__ __forceinline__ void rowTighten(const DevProblem& p, const RowAcc& acc, double lhs, double rhs, long long base,
   int stride, int first, int step, int len, CandQueue& queue, const double (&alpha)[ALPHA_SLOTS])
{
   RowInfo ri;
   ri.acc = acc;
   ri.lhs = lhs;
   ri.rhs = rhs;
   bool cutoff = false;
   if( rowGates(p.num, ri, len, cutoff) )
      rowCandidates<ALPHA, LISTED>(p, ri, base, stride, first, step, len, cutoff, queue, alpha);
   if( cutoff || rowInfeasible(p.num, ri.acc, lhs, rhs) )
      p.ctrl->cutoff = 1;
}

// exact activities of the elements first, first+step, ... < len (four independent loads in flight)
template <bool ALPHA>
__device__ __forceinline__ void accumulateExact(const DevProblem& p, RowAcc& acc, long long base, int stride, int first,
   int step, int len, double (&alpha)[ALPHA_SLOTS])
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
            b[q] = p.bnd[cj[q] & COL_MASK];
      }
#pragma unroll
      for( int q = 0; q < 4; ++q )
      {
         if( k0 + q * step < len )
            accElem(p.num, acc, a[q], b[q].x, b[q].y);
      }
      if( ALPHA && k0 < 8 * step )      // the first two trips (step == 32: k0 == first, first + 128)
      {
#pragma unroll
         for( int q = 0; q < 4; ++q )
         {
            const double al = k0 + q * step < len ? fabs(a[q]) * (b[q].y - b[q].x) : 0.0;
            if( k0 < 4 * step )
               alpha[q] = al;
            else
               alpha[4 + q] = al;
         }
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

// ---- the filter sweep over the CSR stream --------------------------------------------------------------------------

// a marked row whose sums are complete: is it finished (0), or does it need the exact rules (1)?
__device__ __forceinline__ int finishRow(const DevProblem& p, int row, const LeanAcc& tot, unsigned char flag, int lenword,
   const double2& sd, unsigned& nnzdone)
{
   if( flag != ROW_MARKED )
      return 0;
   p.dirty[row] = ROW_CLEAN;
   const bool exact = (lenword & ROWLEN_EXACT) != 0;
   const int len = lenword & ROWLEN_MASK;
   nnzdone += (unsigned)len;
   return (exact || !rowClearlyQuiet(p.num, tot, len, sd.x, sd.y)) ? 1 : 0;
}

// appends `row` of every lane with want == true to a work list (one atomic per warp)
__device__ __forceinline__ void pushRow(const DevProblem& p, bool want, int row, int lane, int list, int listoff)
{
   const unsigned m = __ballot_sync(0xffffffffu, want);
   if( m == 0u )
      return;
   unsigned pos = 0u;
   if( lane == 0 )
      pos = atomicAdd(&p.ctrl->nexact[list], (unsigned)__popc(m));
   pos = __shfl_sync(0xffffffffu, pos, 0);
   if( want )
      p.xlist[listoff + pos + __popc(m & ((1u << lane) - 1u))] = row;
}

// may the sums of the filter stand in for the exact activities of this row?  (see FastAcc)
__device__ __forceinline__ bool fastAccValid(int lenword, bool intbounds, const LeanAcc& la)
{
   return intbounds && (lenword & (ROWLEN_INT | ROWLEN_EXACT)) == ROWLEN_INT && la.cabshi < FAST_CABSHI && la.maxdelta < FAST_MAXDELTA;
}

// the same for the rows of the thread-per-row class: a row with fast == true goes to flist together with its activities
// and its header (length, offset of its slice / 32, sides)
__device__ __forceinline__ void pushShortRow(const DevProblem& p, bool want, bool fast, int row, int lane, const LeanAcc& la,
   int len, int off32, const double2& sd)
{
   const unsigned mg = __ballot_sync(0xffffffffu, want && !fast);
   const unsigned mf = __ballot_sync(0xffffffffu, want && fast);
   if( (mg | mf) == 0u )
      return;
   unsigned posg = 0u;
   unsigned posf = 0u;
   if( lane == 0 )
   {
      if( mg != 0u )
         posg = atomicAdd(&p.ctrl->nexact[0], (unsigned)__popc(mg));
      if( mf != 0u )
         posf = atomicAdd(&p.ctrl->nexact[3], (unsigned)__popc(mf));
   }
   posg = __shfl_sync(0xffffffffu, posg, 0);
   posf = __shfl_sync(0xffffffffu, posf, 0);
   const unsigned below = (1u << lane) - 1u;
   if( want && !fast )
      p.xlist[posg + __popc(mg & below)] = row;
   if( want && fast )
   {
      const unsigned q = posf + __popc(mf & below);
      p.flist[q] = row;
      double2* dst = reinterpret_cast<double2*>(p.facc + q);
      dst[0] = make_double2(la.minact, la.maxact);
      dst[1] = make_double2((double)la.maxdelta, __hiloint2double(off32, len));
      dst[2] = sd;
   }
}

// 32-byte streaming loads (LDG.256): one full sector per lane and request
__device__ __forceinline__ void ldStream256(const double* ptr, double& a, double& b, double& c, double& d)
{
   asm volatile("ld.global.cs.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(ptr));
}
__device__ __forceinline__ void ldStream256(const int* ptr, int (&v)[TPL])
{
   asm volatile("ld.global.cs.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                : "l"(ptr));
}

struct TileRegs
{
   double   a[TPL];
   int      cj[TPL];
   unsigned em;
   int      row0;
};

__device__ __forceinline__ void loadTile(const DevProblem& p, int t, int lane, TileRegs& r)
{
   const long long e0 = p.streambase + (long long)t * TILE + lane * TPL;
   ldStream256(p.vals + e0, r.a[0], r.a[1], r.a[2], r.a[3]);
   ldStream256(p.vals + e0 + 4, r.a[4], r.a[5], r.a[6], r.a[7]);
   ldStream256(p.cols + e0, r.cj);
   r.em = p.endmask[(size_t)t * 32 + lane];
   r.row0 = p.tile_row0[t];
}

__global__ void __launch_bounds__(SWEEP_THREADS, 4) sweep_stream_kernel(const DevProblem p)
{
   TRACE_KERNEL_START(p.ctrl, TR_SWEEP_STREAM);
   static_assert(TPL == 8 && TILE == 256, "the tile layout is wired into the vector loads");
   // per warp: the sums of the rows that end in the current tile, in row order (at most one row per nonzero)
   __shared__ LeanAcc s_tot[SWEEP_THREADS / 32][TILE];

   const int lane = threadIdx.x & 31;
   LeanAcc* tot = s_tot[threadIdx.x >> 5];
   const int gw = (blockIdx.x * SWEEP_THREADS + threadIdx.x) >> 5;
   const int nw = (gridDim.x * SWEEP_THREADS) >> 5;
   // contiguous tile range of this warp inside the tiles [st0, st1) of this rank (a row belongs to the warp -- and the
   // rank -- in whose range it ends: the look-back below reads what lies in front of the range)
   const int per = (p.st1 - p.st0 + nw - 1) / nw;
   const int t0 = p.st0 + gw * per;
   const int t1 = min(p.st1, t0 + per);
   if( t0 >= t1 )
      return;

   unsigned nnzdone = 0;
   LeanAcc carry;                 // sums of the row that is open at the start of the tile (warp-uniform)
   leanInit(carry);
   bool opendirty = false;        // ... and it is marked: the tiles it runs through must be processed

   // look-back: a marked row that started in the range of the previous warp and ends in this one is ours to finish
   {
      const int r = p.tile_row0[t0];
      if( r < p.nsx )
      {
         const long long beg = p.rowbeg[r];
         const long long first = p.streambase + (long long)t0 * TILE;
         if( beg < first && p.dirty[r] == ROW_MARKED )
         {
            LeanAcc part;
            leanInit(part);
            for( long long e = beg + lane; e < first; e += 32 )
            {
               const double2 b = p.bnd[p.cols[e] & COL_MASK];
               leanElem(part, p.vals[e], b.x, b.y);
            }
            leanWarpReduce(part);
            carry = part;
            opendirty = true;
         }
      }
   }

   TileRegs cur;
   bool have = false;             // `cur` already holds the tile that is processed next (prefetched)
   for( int tb = t0; tb < t1; tb += 32 )
   {
      // flags of the next 32 tiles of the range: one load
      const int mine = tb + lane;
      const unsigned flags = __ballot_sync(0xffffffffu, mine < t1 && p.tileflag[mine] != 0);
      if( mine < t1 )
         p.tileflag[mine] = 0;
      for( int i = 0; i < 32 && tb + i < t1; ++i )
      {
         const int t = tb + i;
         if( !(((flags >> i) & 1u) != 0u || opendirty) )
         {
            leanInit(carry);
            continue;
         }
         if( !have )
            loadTile(p, t, lane, cur);
         have = false;

         // ---- bounds of this tile: 8 independent 16-byte gathers per lane
         double2 b[TPL];
#pragma unroll
         for( int k = 0; k < TPL; ++k )
            b[k] = p.bnd[cur.cj[k] & COL_MASK];
         double a[TPL];
#pragma unroll
         for( int k = 0; k < TPL; ++k )
            a[k] = cur.a[k];
         const unsigned em = cur.em;
         const int row0 = cur.row0;

         // ---- the next tile streams in while this one is worked on
         const bool nextflag = i + 1 < 32 && t + 1 < t1 && ((flags >> (i + 1)) & 1u) != 0u;
         if( nextflag )
         {
            loadTile(p, t + 1, lane, cur);
            have = true;
         }

         // the rows that end in this tile are row0, row0+1, ..., row0+nrows-1; this lane ends rows slot0, slot0+1, ...
         const int nends = __popc(em);
         int incl = nends;
#pragma unroll
         for( int d = 1; d < 32; d <<= 1 )
         {
            const int s = __shfl_up_sync(0xffffffffu, incl, d);
            if( lane >= d )
               incl += s;
         }
         const int slot0 = incl - nends;
         const int nrows = __shfl_sync(0xffffffffu, incl, 31);

         // flag, length and sides of row0 + lane: in flight beside the gathers (lane-per-row, coalesced)
         unsigned char rflag = ROW_CLEAN;
         int rlen = 0;
         double2 rsd = make_double2(0.0, 0.0);
         if( lane < nrows )
         {
            rflag = p.dirty[row0 + lane];
            rlen = p.rowlen[row0 + lane];
            rsd = p.sides[row0 + lane];
         }

         // ---- lane-local pass: `head` = sums up to the first row end (to be completed with the lanes before), rows
         // ---- that begin and end inside the lane go straight to their slot, `run` = sums behind the last end
         LeanAcc run;
         LeanAcc head;
         leanInit(run);
         leanInit(head);
         if( !__any_sync(0xffffffffu, nends > 1) )
         {
            // common case, branch free: at most one row ends in this lane, at element `ep`
            const int ep = em != 0u ? __ffs(em) - 1 : TPL;
#pragma unroll
            for( int k = 0; k < TPL; ++k )
            {
               LeanAcc x;
               leanInit(x);
               leanElem(x, a[k], b[k].x, b[k].y);
               if( k <= ep )
                  leanAdd(head, x);
               else
                  leanAdd(run, x);
            }
            if( em == 0u )
            {
               run = head;
               leanInit(head);
            }
         }
         else
         {
            int ends = 0;
#pragma unroll
            for( int k = 0; k < TPL; ++k )
            {
               leanElem(run, a[k], b[k].x, b[k].y);
               if( (em >> k) & 1u )
               {
                  if( ends == 0 )
                     head = run;
                  else
                     tot[slot0 + ends] = run;
                  ++ends;
                  leanInit(run);
               }
            }
         }

         // ---- segmented scan of the tails over the lanes; the open row of the previous tile enters at lane 0
         LeanAcc s = run;
         bool f = (em != 0u);
         if( lane == 0 && !f )
            leanAdd(s, carry);
#pragma unroll
         for( int d = 1; d < 32; d <<= 1 )
         {
            const LeanAcc o = leanShflUp(s, d);
            const bool of = __shfl_up_sync(0xffffffffu, f, d);
            if( lane >= d && !f )
            {
               leanAdd(s, o);
               f = of;
            }
         }
         // exclusive: what the lanes before contribute to the row that ends first in this lane
         LeanAcc pre = leanShflUp(s, 1);
         if( lane == 0 )
            pre = carry;
         if( em != 0u )
         {
            leanAdd(head, pre);
            tot[slot0] = head;
         }
         // the row that is open behind the last end of the tile
         carry = leanShfl(s, 31);
         __syncwarp();

         // ---- one lane per finished row: clearly quiet, or over to the exact kernel
         for( int r0 = 0; r0 < nrows; r0 += 32 )
         {
            const int rr = r0 + lane;
            int verdict = 0;
            if( rr < nrows )
            {
               if( r0 > 0 )
               {
                  rflag = p.dirty[row0 + rr];
                  rlen = p.rowlen[row0 + rr];
                  rsd = p.sides[row0 + rr];
               }
               verdict = finishRow(p, row0 + rr, tot[rr], rflag, rlen, rsd, nnzdone);
            }
            pushRow(p, verdict != 0, row0 + rr, lane, 1, p.nsell);
         }
         __syncwarp();

         // does a marked row run on into the next tile?  (only matters if that tile is not flagged itself)
         opendirty = false;
         if( !nextflag && t + 1 < t1 )
         {
            const int ropen = row0 + nrows;
            const bool lastIsEnd = ((__shfl_sync(0xffffffffu, em, 31) >> (TPL - 1)) & 1u) != 0u;
            if( !lastIsEnd && ropen < p.nsx )
               opendirty = p.dirty[ropen] == ROW_MARKED;
         }
      }
   }
   nnzdone = __reduce_add_sync(0xffffffffu, nnzdone);
   if( lane == 0 )
      addRoundNnz(p, (unsigned long long)nnzdone, gw);
}

// ---- thread-per-row on SELL-32 slices (rows of 1..32 nonzeros) -------------------------------------------------------
// Element k of the row of lane t sits at sell_off[slice] + 32 k + t: every load of a warp is one coalesced line, no
// shuffles at all.  Persistent warps take slices w, w+W, ... in batches of four whose flags / lengths / offsets are
// fetched with one round trip; the coefficients of the next chunk -- of the same or of the next slice -- stream in
// while the bounds of the current chunk are gathered.
constexpr int SELL_THREADS = 256;
constexpr int SELL_NB = 4;

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

template <int CH>
__device__ __forceinline__ void sellSweep(const DevProblem& p, int nblockthreads, bool intbounds, int sbeg, int send, unsigned& nnzdone)
{
   const Num& n = p.num;
   const int lane = threadIdx.x & 31;
   const int gw = (blockIdx.x * nblockthreads + threadIdx.x) >> 5;
   const int nw = (gridDim.x * nblockthreads) >> 5;
   // this rank's slices of [sbeg, send): sbeg + rank, sbeg + rank + nranks, ... = slice number q of the share
   const int nq = (send - sbeg - p.rank + p.nranks - 1) / p.nranks;
   for( int q0 = gw; q0 < nq; q0 += SELL_NB * nw )
   {
      // ---- one round trip: flags, lengths and offsets of the next four slices of this warp
      int len[SELL_NB];
      int maxlen[SELL_NB];
      long long base[SELL_NB];
      unsigned actm = 0u;
      unsigned exactm = 0u;
      unsigned intm = 0u;
#pragma unroll
      for( int i = 0; i < SELL_NB; ++i )
      {
         const int q = q0 + i * nw;
         const int slice = sbeg + p.rank + q * p.nranks;
         const int row = slice * 32 + lane;
         const bool valid = q < nq && row < p.nsell;
         const unsigned char f = valid ? p.dirty[row] : ROW_CLEAN;
         const int lw = valid ? p.rowlen[row] : 0;
         base[i] = (q < nq ? p.sell_off[slice] : 0) + lane;
         const bool act = f == ROW_MARKED;
         len[i] = act ? (lw & ROWLEN_MASK) : 0;
         if( act )
            actm |= 1u << i;
         if( act && (lw & ROWLEN_EXACT) != 0 )
            exactm |= 1u << i;
         if( act && (lw & ROWLEN_INT) != 0 )
            intm |= 1u << i;
      }
#pragma unroll
      for( int i = 0; i < SELL_NB; ++i )
         maxlen[i] = __reduce_max_sync(0xffffffffu, len[i]);

      double an[CH];
      int cjn[CH];
      loadChunk<CH>(p, base[0], 0, len[0], an, cjn);
#pragma unroll
      for( int i = 0; i < SELL_NB; ++i )
      {
         const int row = (sbeg + p.rank + (q0 + i * nw) * p.nranks) * 32 + lane;
         const bool act = ((actm >> i) & 1u) != 0u;
         double2 sd = make_double2(0.0, 0.0);
         if( act )
            sd = p.sides[row];
         LeanAcc acc;
         leanInit(acc);
         for( int c = 0; c < maxlen[i]; c += CH )
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
            if( c + CH < maxlen[i] )
               loadChunk<CH>(p, base[i], c + CH, len[i], an, cjn);
            else if( i + 1 < SELL_NB )
               loadChunk<CH>(p, base[i + 1], 0, len[i + 1], an, cjn);
#pragma unroll
            for( int k = 0; k < CH; ++k )
            {
               if( c + k < len[i] )
               {
                  b[k] = p.bnd[cj[k] & COL_MASK];
               }
            }
#pragma unroll
            for( int k = 0; k < CH; ++k )
            {
               if( c + k < len[i] )
                  leanElem(acc, a[k], b[k].x, b[k].y);
            }
         }
         if( maxlen[i] == 0 && i + 1 < SELL_NB )
            loadChunk<CH>(p, base[i + 1], 0, len[i + 1], an, cjn);
         bool handoff = false;
         bool fast = false;
         if( act )
         {
            const bool exact = ((exactm >> i) & 1u) != 0u;
            handoff = exact || !rowClearlyQuiet(n, acc, len[i], sd.x, sd.y);
            fast = handoff && fastAccValid(exact ? ROWLEN_EXACT : (((intm >> i) & 1u) != 0u ? ROWLEN_INT : 0), intbounds, acc);
            p.dirty[row] = ROW_CLEAN;
            nnzdone += (unsigned)len[i];
         }
         pushShortRow(p, handoff, fast, row, lane, acc, len[i], (int)((base[i] - lane) >> 5), sd);
      }
   }
}

template <int CH, int MINB>
__global__ void __launch_bounds__(SELL_THREADS, MINB) sweep_sell_kernel(const DevProblem p)
{
   TRACE_KERNEL_START(p.ctrl, TR_SWEEP_SELL);
   unsigned nnzdone = 0;
   const bool intbounds = *p.fracflag == 0u;
   sellSweep<CH>(p, SELL_THREADS, intbounds, 0, p.nsellunit >> 5, nnzdone);
   sellSweep<CH>(p, SELL_THREADS, intbounds, p.nsellunit >> 5, (p.nsell + 31) >> 5, nnzdone);
   nnzdone = __reduce_add_sync(0xffffffffu, nnzdone);
   if( (threadIdx.x & 31) == 0 )
      addRoundNnz(p, (unsigned long long)nnzdone, (blockIdx.x * SELL_THREADS + threadIdx.x) >> 5);
}

// ---- the SELL sweep with a shared-memory bit table -------------------------------------------------------------------
// The gather variant above is bound by the L1/L2 path, not by HBM: a random 16-byte bound gather costs one L1 wavefront
// and one L2 sector per nonzero (10M gathers >= 40 us on C3).  Most columns of a MIP are binaries at their original
// bounds: freebits holds one bit per column, "the bounds are exactly (0,1)", kept up to date by everybody who writes bnd
// (noteBounds).  One block per SM stages the table into shared memory with bulk copies (TMA, 125 KB for 1M columns;
// columns beyond nfreecols and columns whose bit is clear are gathered as before): 32 random words cost ~4 bank-conflict
// cycles per warp instead of 32 wavefronts.  (A variant that also staged the matrix through per-warp cp.async / TMA rings
// in the remaining shared memory was tried and was slower: with the table resident, ~100 KB of ring per SM do not hold
// more bytes in flight than the registers do, and the ring bookkeeping costs more instructions than it saves latency.)
constexpr int SB_MAXSMEM = 227 * 1024;
constexpr int SB_AUX_BYTES = 256;                          // the mbarrier of the table + alignment
constexpr int SELLBITS_MAXWORDS = ((SB_MAXSMEM - SB_AUX_BYTES) / 4) & ~31;

__device__ __forceinline__ unsigned smemAddr(const void* ptr)
{
   return (unsigned)__cvta_generic_to_shared(ptr);
}
__device__ __forceinline__ void mbarInit(unsigned bar, unsigned count)
{
   asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbarExpectTx(unsigned bar, unsigned bytes)
{
   asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbarWait(unsigned bar, unsigned parity)
{
   unsigned ok;
   do
   {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
   } while( ok == 0u );
}
// global -> shared bulk copy through the async proxy (TMA, 1-D); bytes and both addresses are multiples of 16
__device__ __forceinline__ void bulkLoad(unsigned dst, const void* src, unsigned bytes, unsigned bar)
{
   asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// The filter sums in midpoint / half-width form: with m_k = a_k (l_k + u_k)/2 and h_k = |a_k| (u_k - l_k)/2 the term of
// the minimal activity is m_k - h_k and that of the maximal activity m_k + h_k, whatever the sign of a_k -- no selects:
//    minact = M - H,  maxact = M + H,  maxdelta = 2 max_k h_k,   M = sum m_k,  H = sum h_k           (7 instructions per nonzero)
// The table bndf holds ((l+u)/2, (u-l)/2) per column.  Error against the exact rules: every m_k, h_k carries at most three
// roundings, a sum of n terms n - 1 more, so each activity is within (n + 4) u C of the exact value, C = sum |m_k| + h_k
// (tracked as Mabs + H, an upper bound of every |a_k l_k|, |a_k u_k| as well) -- covered by rowClearlyQuiet's
// E = (n^2 + 8) 2.3e-16 cabs with cabs >= C.  An infinite or NaN bound makes C huge or NaN: not quiet.
struct MidAcc
{
   double M;
   double Mabs;
   double H;
   float  hmax;     // max_k h_k, rounded up
};

__device__ __forceinline__ void midInit(MidAcc& r)
{
   r.M = r.Mabs = r.H = 0.0;
   r.hmax = 0.0f;
}

__device__ __forceinline__ void midElem(MidAcc& r, double a, double mid, double hw)
{
   const double m = a * mid;
   const double h = fabs(a) * hw;
   r.M += m;
   r.Mabs += fabs(m);
   r.H += h;
   r.hmax = fmaxf(r.hmax, __double2float_ru(h));
}

__device__ __forceinline__ LeanAcc midToLean(const MidAcc& r)
{
   LeanAcc o;
   o.minact = r.M - r.H;
   o.maxact = r.M + r.H;
   o.maxdelta = __fmul_ru(2.0f, r.hmax);
   o.cabshi = __double2hiint(r.Mabs + r.H) & 0x7fffffff;
   return o;
}

// bound pair of a column unless `skip` (a predicated load: no branch, no reconvergence point)
__device__ __forceinline__ void gatherUnless(unsigned skip, const double2* addr, double2& b)
{
   asm("{\n\t.reg .pred q;\n\tsetp.eq.u32 q, %2, 0;\n\t@q ld.global.v2.f64 {%0, %1}, [%3];\n\t}"
       : "+d"(b.x), "+d"(b.y) : "r"(skip), "l"(addr));
}

// coefficient of a nonzero of a unit row (+1 or -1) from its sign word: bit 30 of the column word (COL_NEGCOEF) moved to
// bit 31, the position of the sign in the high word of a double
__device__ __forceinline__ int unitSign(int colword)
{
   return (colword << 1) & (int)0x80000000u;
}
// one chunk of CH nonzeros of a row of a SELL slice; UNIT: the column words only (the values are not read)
template <int CH, bool UNIT>
__device__ __forceinline__ void loadChunkBits(const DevProblem& p, long long base, int c, int len, double (&a)[CH], int (&cj)[CH])
{
#pragma unroll
   for( int k = 0; k < CH; ++k )
   {
      if( c + k < len )
      {
         if( !UNIT )
            a[k] = ldStream(p.vals + base + 32LL * (c + k));
         cj[k] = ldStream(p.cols + base + 32LL * (c + k));
      }
   }
}

// the slices [sbeg, send) of the SELL bin, thread per row; sums in midpoint / half-width form from bndf
//      ALLCOLS: the table covers every column (no range checks)
//      UNIT: every coefficient of these slices is +1 or -1 (rows [0, nsellunit)): 4 instead of 12 bytes per nonzero
template <int NT, int CH, bool ALLCOLS, bool UNIT>
__device__ __forceinline__ void sellBitsRange(const DevProblem& p, const unsigned* s_free, unsigned tabbar, bool& tabready,
   bool intbounds, int sbeg, int send, unsigned& nnzdone)
{
   const Num& n = p.num;
   const int lane = threadIdx.x & 31;
   // warp-major numbering: warp w of block b is warp w * gridDim + b.  The slices do not divide evenly among the warps
   // (C3: 6.6 each); with block-major numbering the blocks with the low numbers would hold all the warps that take
   // one slice more (and of the longest rows), and the other SMs would idle for the last seventh of the kernel
   const int gw = (int)(threadIdx.x >> 5) * (int)gridDim.x + (int)blockIdx.x;
   const int nw = (gridDim.x * NT) >> 5;
   const int lastword = p.nfreewords - 1;
   // (taking the slices from the end of the range, longest rows first, so that the extra slice of some warps is a short one,
   // was measured: 45 instead of 41 us)
   // this rank's slices of [sbeg, send): sbeg + rank, sbeg + rank + nranks, ... = slice number q of the share
   const int nq = (send - sbeg - p.rank + p.nranks - 1) / p.nranks;
   for( int q0 = gw; q0 < nq; q0 += SELL_NB * nw )
   {
      // ---- one round trip: flags, lengths and offsets of the next four slices of this warp
      int len[SELL_NB];
      int maxlen[SELL_NB];
      long long base[SELL_NB];
      unsigned actm = 0u;
      unsigned exactm = 0u;
      unsigned intm = 0u;
#pragma unroll
      for( int i = 0; i < SELL_NB; ++i )
      {
         const int q = q0 + i * nw;
         const int slice = sbeg + p.rank + q * p.nranks;
         const int row = slice * 32 + lane;
         const bool valid = q < nq && row < p.nsell;
         const unsigned char f = valid ? p.dirty[row] : ROW_CLEAN;
         const int lw = valid ? p.rowlen[row] : 0;
         base[i] = (q < nq ? p.sell_off[slice] : 0) + lane;
         const bool act = f == ROW_MARKED;
         len[i] = act ? (lw & ROWLEN_MASK) : 0;
         if( act )
            actm |= 1u << i;
         if( act && (lw & ROWLEN_EXACT) != 0 )
            exactm |= 1u << i;
         if( act && (lw & ROWLEN_INT) != 0 )
            intm |= 1u << i;
      }
#pragma unroll
      for( int i = 0; i < SELL_NB; ++i )
         maxlen[i] = __reduce_max_sync(0xffffffffu, len[i]);

      double an[CH];
      int cjn[CH];
      loadChunkBits<CH, UNIT>(p, base[0], 0, len[0], an, cjn);
      if( !tabready )
      {
         mbarWait(tabbar, 0u);      // the first coefficients are on their way while the table arrives
         tabready = true;
      }
#pragma unroll
      for( int i = 0; i < SELL_NB; ++i )
      {
         const int row = (sbeg + p.rank + (q0 + i * nw) * p.nranks) * 32 + lane;
         const bool act = ((actm >> i) & 1u) != 0u;
         double2 sd = make_double2(0.0, 0.0);
         if( act )
            sd = p.sides[row];
         MidAcc macc;
         midInit(macc);
         for( int c = 0; c < maxlen[i]; c += CH )
         {
            double a[CH];
            int sg[CH];      // UNIT: the sign bit of the coefficient (+1 / -1), in the position of a double's high word
            int cj[CH];
            double2 b[CH];
#pragma unroll
            for( int k = 0; k < CH; ++k )
            {
               if( UNIT )
                  sg[k] = unitSign(cjn[k]);
               else
                  a[k] = an[k];
               cj[k] = cjn[k] & COL_MASK;
            }
            if( c + CH < maxlen[i] )
               loadChunkBits<CH, UNIT>(p, base[i], c + CH, len[i], an, cjn);
            else if( i + 1 < SELL_NB )
               loadChunkBits<CH, UNIT>(p, base[i + 1], 0, len[i + 1], an, cjn);
#pragma unroll
            for( int k = 0; k < CH; ++k )
            {
               if( c + k < len[i] )
               {
                  const unsigned w = s_free[ALLCOLS ? cj[k] >> 5 : min(cj[k] >> 5, lastword)];
                  const unsigned fr = (ALLCOLS || cj[k] < p.nfreecols) ? (w >> (cj[k] & 31)) & 1u : 0u;
                  b[k] = make_double2(0.5, 0.5);
                  gatherUnless(fr, p.bndf + cj[k], b[k]);
               }
            }
#pragma unroll
            for( int k = 0; k < CH; ++k )
            {
               if( c + k < len[i] )
               {
                  if( UNIT )
                  {
                     // a = +1 / -1: a * mid is mid with the sign flipped, |a| * hw is hw -- the same values as midElem
                     // would compute with a = +1 / -1, without the two multiplications
                     const double m = __hiloint2double(__double2hiint(b[k].x) ^ sg[k], __double2loint(b[k].x));
                     macc.M += m;
                     macc.Mabs += fabs(b[k].x);
                     macc.H += b[k].y;
                     macc.hmax = fmaxf(macc.hmax, __double2float_ru(b[k].y));
                  }
                  else
                     midElem(macc, a[k], b[k].x, b[k].y);
               }
            }
         }
         if( maxlen[i] == 0 && i + 1 < SELL_NB )
            loadChunkBits<CH, UNIT>(p, base[i + 1], 0, len[i + 1], an, cjn);
         bool handoff = false;
         bool fast = false;
         const LeanAcc la = midToLean(macc);
         if( act )
         {
            const bool exact = ((exactm >> i) & 1u) != 0u;
            handoff = exact || !rowClearlyQuiet(n, la, len[i], sd.x, sd.y);
            fast = handoff && fastAccValid(exact ? ROWLEN_EXACT : (((intm >> i) & 1u) != 0u ? ROWLEN_INT : 0), intbounds, la);
            p.dirty[row] = ROW_CLEAN;
            nnzdone += (unsigned)len[i];
         }
         pushShortRow(p, handoff, fast, row, lane, la, len[i], (int)((base[i] - lane) >> 5), sd);
      }
   }
}

// ---- the unit slices (rows [0, nsellunit): every coefficient is +1 or -1) -------------------------------------------------
// A column whose bit is set in the table has the bounds (0,1): its term is m = +-1/2, h = 1/2.  Such nonzeros are COUNTED
// (the free columns, and those of them with a negative coefficient, as two 16-bit fields of one register) instead of summed
// in fp64: three integer instructions, no constants to materialise; the sums of the counted part, (nf - 2 nn)/2 and nf/2,
// are exact.  A column that is not free is looked up in colstate[] (one byte: eight lookups in flight cost eight
// registers, eight bound pairs would cost 32): fixed at 0 -- no term at all; fixed at 1 -- counted like the free ones
// (m = +-1, h = 0); only what is left (general integers, continuous columns) takes the fp64 path with its gather of bndf.
// Both steps sit behind warp votes: a chunk whose columns are all free never sees the second, a chunk of binaries never
// the third.
// Measured (C3, 10M nonzeros, full round): 38.9 us against 40.3 us for the version that summed every nonzero in fp64; chunks
// of 8 or 16 nonzeros, 768 / 512 threads with more registers, and flags / lengths / offsets fetched a batch ahead were all
// slower (53 - 104 us: the 64-register budget of a 1024-thread block spills, and a spilled load is waited for at once).
constexpr int UNIT_NB = 2;           // slices per batch

// FULL: all nonzeros of the chunk exist in every lane (a slice of rows of one length, all marked): no predicates
template <int UNIT_CH, bool ALLCOLS, bool FULL>
__device__ __forceinline__ void unitChunk(const DevProblem& p, const unsigned* s_free, const int (&w)[UNIT_CH], int nvalid,
   unsigned& cntfree, unsigned& cntone, MidAcc& macc)
{
   unsigned notfree = 0u;
#pragma unroll
   for( int k = 0; k < UNIT_CH; ++k )
   {
      const bool there = FULL || k < nvalid;                      // (a word that was not loaded must not index the table)
      const unsigned wk = there ? (unsigned)w[k] : 0u;
      const unsigned col = wk & (unsigned)COL_MASK;
      const unsigned word = s_free[ALLCOLS ? col >> 5 : min(col >> 5, (unsigned)(p.nfreewords - 1))];
      unsigned fr = (word >> (col & 31u)) & 1u;
      if( !ALLCOLS && col >= (unsigned)p.nfreecols )
         fr = 0u;
      if( !there )
         fr = 0u;
      // low half: free columns; high half: free columns with a negative coefficient (COL_NEGCOEF is bit 30 of the word)
      cntfree += fr * (1u + ((wk >> 14) & 0x10000u));
      if( there && fr == 0u )
         notfree |= 1u << k;
   }
   if( __any_sync(0xffffffffu, notfree != 0u) )
   {
      unsigned code[UNIT_CH];      // (not bytes: a byte array would live in local memory)
#pragma unroll
      for( int k = 0; k < UNIT_CH; ++k )
      {
         if( (notfree >> k) & 1u )
            code[k] = p.colstate[w[k] & COL_MASK];
      }
      unsigned other = 0u;
#pragma unroll
      for( int k = 0; k < UNIT_CH; ++k )
      {
         if( (notfree >> k) & 1u )
         {
            if( code[k] == CS_FIX1 )
               cntone += 1u + (((unsigned)w[k] >> 14) & 0x10000u);
            else if( code[k] != CS_FIX0 )
               other |= 1u << k;
         }
      }
      if( __any_sync(0xffffffffu, other != 0u) )
      {
#pragma unroll
         for( int k = 0; k < UNIT_CH; ++k )
         {
            if( (other >> k) & 1u )
            {
               // a = +1 / -1: a * mid is mid with the sign flipped, |a| * hw is hw (the values midElem would compute)
               const double2 b = p.bndf[w[k] & COL_MASK];
               const double m = __hiloint2double(__double2hiint(b.x) ^ unitSign(w[k]), __double2loint(b.x));
               macc.M += m;
               macc.Mabs += fabs(b.x);
               macc.H += b.y;
               macc.hmax = fmaxf(macc.hmax, __double2float_ru(b.y));
            }
         }
      }
   }
}

template <int UNIT_CH>
__device__ __forceinline__ void loadUnitChunk(const DevProblem& p, long long base, int c, int len, int (&w)[UNIT_CH])
{
#pragma unroll
   for( int k = 0; k < UNIT_CH; ++k )
   {
      if( c + k < len )
         w[k] = ldStream(p.cols + base + 32LL * (c + k));
   }
}

// flags, length words and offsets of the slices of one batch of a warp
struct UnitHdr
{
   int           lw[UNIT_NB];
   int           off32[UNIT_NB];
   unsigned char f[UNIT_NB];
};

template <int NT, bool ALLCOLS, int UNIT_CH>
__device__ __forceinline__ void sellUnitRange(const DevProblem& p, const unsigned* s_free, unsigned tabbar, bool& tabready,
   bool intbounds, int sbeg, int send, unsigned& nnzdone)
{
   const Num& n = p.num;
   const int lane = threadIdx.x & 31;
   const int gw = (int)(threadIdx.x >> 5) * (int)gridDim.x + (int)blockIdx.x;      // warp-major, see sellBitsRange
   const int nw = (gridDim.x * NT) >> 5;
   const int nq = (send - sbeg - p.rank + p.nranks - 1) / p.nranks;
   auto loadHdr = [&](int q0, UnitHdr& h)
   {
#pragma unroll
      for( int i = 0; i < UNIT_NB; ++i )
      {
         const int q = q0 + i * nw;
         const int slice = sbeg + p.rank + q * p.nranks;
         const int row = slice * 32 + lane;
         const bool valid = q < nq && row < p.nsell;
         h.f[i] = valid ? p.dirty[row] : ROW_CLEAN;
         h.lw[i] = valid ? p.rowlen[row] : 0;
         h.off32[i] = q < nq ? p.sell_off32[slice] : 0;
      }
   };
   int wn[UNIT_CH];
   for( int q0 = gw; q0 < nq; q0 += UNIT_NB * nw )
   {
      UnitHdr h;
      loadHdr(q0, h);
      int len[UNIT_NB];
      int maxlen[UNIT_NB];
      bool uniform[UNIT_NB];
#pragma unroll
      for( int i = 0; i < UNIT_NB; ++i )
      {
         len[i] = h.f[i] == ROW_MARKED ? (h.lw[i] & ROWLEN_MASK) : 0;
         maxlen[i] = __reduce_max_sync(0xffffffffu, len[i]);
         uniform[i] = __reduce_min_sync(0xffffffffu, len[i]) == maxlen[i];
      }
      loadUnitChunk<UNIT_CH>(p, 32LL * h.off32[0] + lane, 0, len[0], wn);
      if( !tabready )
      {
         mbarWait(tabbar, 0u);      // the first column words are on their way while the table arrives
         tabready = true;
      }
#pragma unroll
      for( int i = 0; i < UNIT_NB; ++i )
      {
         const int row = (sbeg + p.rank + (q0 + i * nw) * p.nranks) * 32 + lane;
         const long long base = 32LL * h.off32[i] + lane;
         const bool act = h.f[i] == ROW_MARKED;
         double2 sd = make_double2(0.0, 0.0);
         if( act )
            sd = p.sides[row];
         MidAcc macc;
         midInit(macc);
         unsigned cntfree = 0u;
         unsigned cntone = 0u;
         for( int c = 0; c < maxlen[i]; c += UNIT_CH )
         {
            int w[UNIT_CH];
#pragma unroll
            for( int k = 0; k < UNIT_CH; ++k )
               w[k] = wn[k];
            // the chunk after this one
            if( c + UNIT_CH < maxlen[i] )
               loadUnitChunk<UNIT_CH>(p, base, c + UNIT_CH, len[i], wn);
            else if( i + 1 < UNIT_NB )
               loadUnitChunk<UNIT_CH>(p, 32LL * h.off32[i + 1 < UNIT_NB ? i + 1 : i] + lane, 0, len[i + 1 < UNIT_NB ? i + 1 : i], wn);
            if( uniform[i] && c + UNIT_CH <= maxlen[i] )
               unitChunk<UNIT_CH, ALLCOLS, true>(p, s_free, w, UNIT_CH, cntfree, cntone, macc);
            else
               unitChunk<UNIT_CH, ALLCOLS, false>(p, s_free, w, len[i] - c, cntfree, cntone, macc);
         }
         if( maxlen[i] == 0 && i + 1 < UNIT_NB )
            loadUnitChunk<UNIT_CH>(p, 32LL * h.off32[i + 1 < UNIT_NB ? i + 1 : i] + lane, 0, len[i + 1 < UNIT_NB ? i + 1 : i], wn);
         bool handoff = false;
         bool fast = false;
         LeanAcc la;
         leanInit(la);
         if( act )
         {
            // the counted part: nf free columns (nfn of them with coefficient -1), n1 columns fixed at 1 (n1n with -1)
            const int nf = (int)(cntfree & 0xffffu);
            const int nfn = (int)(cntfree >> 16);
            const int n1 = (int)(cntone & 0xffffu);
            const int n1n = (int)(cntone >> 16);
            const double half = 0.5 * (double)nf;
            macc.M += (half - (double)nfn) + (double)(n1 - 2 * n1n);
            macc.Mabs += half + (double)n1;
            macc.H += half;
            if( nf > 0 )
               macc.hmax = fmaxf(macc.hmax, 0.5f);
            la = midToLean(macc);
            handoff = (h.lw[i] & ROWLEN_EXACT) != 0 || !rowClearlyQuiet(n, la, len[i], sd.x, sd.y);
            fast = handoff && fastAccValid(h.lw[i], intbounds, la);
            p.dirty[row] = ROW_CLEAN;
            nnzdone += (unsigned)len[i];
         }
         pushShortRow(p, handoff, fast, row, lane, la, len[i], h.off32[i], sd);
      }
   }
}

// CHU: nonzeros per thread and chunk in the unit slices (sellUnitRange)
template <int NT, int CH, bool ALLCOLS, int CHU>
__global__ void __launch_bounds__(NT, 1) sweep_sell_bits_kernel(const DevProblem p)
{
   TRACE_KERNEL_START(p.ctrl, TR_SWEEP_SELL);
   extern __shared__ __align__(128) unsigned char s_raw[];
   const unsigned tabbytes = (unsigned)p.nfreewords * 4u;
   const unsigned* s_free = reinterpret_cast<const unsigned*>(s_raw);
   const unsigned tabbar = smemAddr(s_raw + ((tabbytes + 127u) & ~127u));
   if( threadIdx.x == 0 )
   {
      mbarInit(tabbar, 1u);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      mbarExpectTx(tabbar, tabbytes);
      for( unsigned o = 0; o < tabbytes; o += 16384u )
         bulkLoad(smemAddr(s_raw) + o, reinterpret_cast<const unsigned char*>(p.freebits) + o, min(16384u, tabbytes - o), tabbar);
   }
   __syncthreads();

   const int lane = threadIdx.x & 31;
   const int gw = (blockIdx.x * NT + threadIdx.x) >> 5;
   unsigned nnzdone = 0;
   bool tabready = false;
   const bool intbounds = *p.fracflag == 0u;
   sellUnitRange<NT, ALLCOLS, CHU>(p, s_free, tabbar, tabready, intbounds, 0, p.nsellunit >> 5, nnzdone);
   sellBitsRange<NT, CH, ALLCOLS, false>(p, s_free, tabbar, tabready, intbounds, p.nsellunit >> 5, (p.nsell + 31) >> 5, nnzdone);
   if( !tabready && threadIdx.x < 32 )
      mbarWait(tabbar, 0u);         // the block must not retire under the copies it issued
   nnzdone = __reduce_add_sync(0xffffffffu, nnzdone);
   if( lane == 0 )
      addRoundNnz(p, (unsigned long long)nnzdone, gw);
}

// ---- block-per-row for rows longer than STREAM_MAXLEN (and empty rows) ---------------------------------------------
__global__ void __launch_bounds__(LONG_THREADS) sweep_long_kernel(const DevProblem p)
{
   TRACE_KERNEL_START(p.ctrl, TR_SWEEP_LONG);
   __shared__ LeanAcc s_lean[LONG_THREADS / 32];

   const int lane = threadIdx.x & 31;
   const int warp = threadIdx.x >> 5;
   const Num& n = p.num;
   const int row0 = p.nsx;
   const int nrows = p.nrows - row0;

   for( int r = p.rank + p.nranks * (int)blockIdx.x; r < nrows; r += p.nranks * (int)gridDim.x )
   {
      const int row = row0 + r;
      __syncthreads();                  // previous row done with the shared state and its flag
      if( p.dirty[row] != ROW_MARKED )  // block-uniform: nobody rewrites the flag before the barrier below
         continue;
      int len = p.rowlen[row];
      const bool exact = (len & ROWLEN_EXACT) != 0;
      len &= ROWLEN_MASK;
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
               b[k] = p.bnd[cj[k] & COL_MASK];
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
            leanAdd(la, s_lean[w]);
         const bool handoff = exact || !rowClearlyQuiet(n, la, len, sd.x, sd.y);
         p.dirty[row] = ROW_CLEAN;
         addRoundNnz(p, (unsigned long long)len, row);
         if( handoff )
            p.xlist[p.nsx + atomicAdd(&p.ctrl->nexact[2], 1u)] = row;
      }
   }
}

// ---- the exact rules for every row the filter sweeps could not finish ---------------------------------------------
// one launch over the three work lists: short rows by groups of 8 lanes (the whole row in registers: activities by a
// butterfly over the group, candidates straight from the registers), medium rows by warps, long rows by blocks
constexpr int EXACT_THREADS = 256;
constexpr int EXACT_G = 8;                           // lanes per short row
constexpr int EXACT_Q = SHORT_MAXLEN / EXACT_G;      // elements per lane

// claims a marked row (sparse rounds: the lists may name a row twice, or a row that a dense sweep finished already)
__device__ __forceinline__ bool claimRow(const DevProblem& p, int row)
{
   unsigned* word = reinterpret_cast<unsigned*>(p.dirty + (row & ~3));
   const int shift = 8 * (row & 3);
   const unsigned old = atomicAnd(word, ~(0xffu << shift));
   return ((old >> shift) & 0xffu) == ROW_MARKED;
}

// SPARSE = false: the rows the filter sweeps handed over (xlist);  SPARSE = true: the rows the last apply step marked
// (marklist) -- in a round with few marked rows the filter pass is skipped and every marked row gets the exact rules
template <bool SPARSE>
__device__ __forceinline__ void exactPhase(const DevProblem& p, const int* list0, unsigned n0, const int* list1, unsigned n1,
   const int* list2, unsigned n2, RowAcc* s_acc, CandQueue* s_queue, int nblockthreads, const int* list0b = nullptr, unsigned n0b = 0u)
{
   const Num& n = p.num;
   const int lane = threadIdx.x & 31;
   const int warp = threadIdx.x >> 5;
   const int gtid = blockIdx.x * nblockthreads + threadIdx.x;
   const int nthreads = gridDim.x * nblockthreads;
   unsigned long long nnzdone = 0;

   // ---- short rows: EXACT_G lanes per row
   {
      const int gl = lane & (EXACT_G - 1);
      const int ngroups = nthreads / EXACT_G;
      // (a second list of short rows: the rows that came with their activities when there are too few of them for the
      // thread-per-row phase, see exact_rows_kernel)
      const unsigned rounds = (n0 + n0b + ngroups - 1) / ngroups;      // warp-uniform trip count
      for( unsigned it = 0; it < rounds; ++it )
      {
         const unsigned item = it * ngroups + gtid / EXACT_G;
         bool valid = item < n0 + n0b;
         int row = 0;
         if( valid )
            row = item < n0 ? list0[item] : list0b[item - n0];
         int len = 0;
         long long base = 0;
         double2 sd = make_double2(0.0, 0.0);
         if( valid )
         {
            // (the header of the row is on its way while the claim below is answered: the list names real rows)
            len = p.rowlen[row] & ROWLEN_MASK;
            base = p.sell_off[row >> 5] + (row & 31);
            sd = p.sides[row];
         }
         if( SPARSE )
         {
            int mine = (valid && gl == 0) ? (claimRow(p, row) ? 1 : 0) : 0;
            mine = __shfl_sync(0xffffffffu, mine, lane & ~(EXACT_G - 1));
            valid = mine != 0;      // (the loads below do not wait for this: a row that is not ours costs a few loads)
            if( valid && gl == 0 )
               nnzdone += (unsigned long long)len;
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
               b[q] = p.bnd[cj[q] & COL_MASK];
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
               sk.colbits = p.colbits;
               sk.chglist = p.chglist;
               sk.nchgcols = &p.ctrl->nchgcols;
               sk.listed = SPARSE;
               static_assert(EXACT_Q == 4, "the slot selection below is written for four slots");
               unsigned pass = 0u;
#pragma unroll
               for( int q = 0; q < EXACT_Q; ++q )
               {
                  if( gl + EXACT_G * q < len && (!ri.easy || passesSlackTest(ri, fabs(a[q]) * (b[q].y - b[q].x), thr)) )
                     pass |= 1u << q;
               }
               while( pass != 0u )      // one copy of the candidate rules for the four slots (see rowCandidates)
               {
                  const int q = __ffs(pass) - 1;
                  pass &= pass - 1u;
                  double aq = a[0];
                  int cq = cj[0];
                  double2 bq = b[0];
                  if( q == 1 ) { aq = a[1]; cq = cj[1]; bq = b[1]; }
                  if( q == 2 ) { aq = a[2]; cq = cj[2]; bq = b[2]; }
                  if( q == 3 ) { aq = a[3]; cq = cj[3]; bq = b[3]; }
                  bool touched = false;
                  candidates(n, sk, ri, aq, cq & COL_MASK, cq < 0, bq.x, bq.y, cutoff, touched);
                  const bool firsttouch = touched && raiseColumnBit(sk, cq & COL_MASK);
                  listChangedColumn(sk, cq & COL_MASK, firsttouch);
               }
            }
            if( cutoff || (gl == 0 && rowInfeasible(n, ri.acc, ri.lhs, ri.rhs)) )
               p.ctrl->cutoff = 1;
         }
      }
   }

   // ---- medium rows: one warp per row
   // (Instruction bound, not latency bound: ncu on C4's first round -- 200k rows of ~250 nonzeros, all of which tighten
   // something -- counts ~2000 warp instructions per row, 45 % issue utilisation at 16 warps per SM, 1.8 long-scoreboard
   // stalls per issue; most of them are the candidate rules, which a good part of a row's nonzeros reach.  Two variants
   // were measured and dropped: this loop as its own kernel with the header and the first 128 column words of the NEXT
   // row and the next 128 of the current row in flight -- 661 us for these rows against ~620 us here; and a candidate
   // queue carried across the rows of a warp, so that the rules always run with 32 entries -- 881 us for the kernel
   // against 765 us: the per-row queue is nearly full anyway, and every entry had to fetch its row's state.)
   {
      const int gw = gtid >> 5;
      const int nw = nthreads >> 5;
      for( unsigned item = gw; item < n1; item += nw )
      {
         const int row = list1[item];
         if( SPARSE )
         {
            int mine = lane == 0 ? (claimRow(p, row) ? 1 : 0) : 0;
            mine = __shfl_sync(0xffffffffu, mine, 0);
            if( !mine )
               continue;
         }
         const int len = p.rowlen[row] & ROWLEN_MASK;
         const long long beg = p.rowbeg[row];
         const double2 sd = p.sides[row];
         if( SPARSE && lane == 0 )
            nnzdone += (unsigned long long)len;
         RowAcc acc;
         accInit(acc);
         double alpha[ALPHA_SLOTS];
#pragma unroll
         for( int q = 0; q < ALPHA_SLOTS; ++q )
            alpha[q] = 0.0;
         accumulateExact<true>(p, acc, beg, 1, lane, 32, len, alpha);
         accWarpReduce(acc, lane);
         rowTighten<true, SPARSE>(p, acc, sd.x, sd.y, beg, 1, lane, 32, len, s_queue[warp], alpha);
      }
   }

   // ---- long rows: one block per row
   for( unsigned item = blockIdx.x; item < n2; item += gridDim.x )
   {
      const int row = list2[item];
      __syncthreads();                  // the previous row's shared state has been read by everybody
      if( SPARSE )
      {
         __shared__ int s_mine;
         if( threadIdx.x == 0 )
            s_mine = claimRow(p, row) ? 1 : 0;
         __syncthreads();
         if( !s_mine )
            continue;
      }
      const int len = p.rowlen[row] & ROWLEN_MASK;
      const long long beg = p.rowbeg[row];
      const double2 sd = p.sides[row];
      if( SPARSE && threadIdx.x == 0 )
         nnzdone += (unsigned long long)len;
      RowAcc acc;
      accInit(acc);
      double noalpha[ALPHA_SLOTS];
      accumulateExact<false>(p, acc, beg, 1, threadIdx.x, nblockthreads, len, noalpha);
      accWarpReduce(acc, lane);
      if( lane == 0 )
         s_acc[warp] = acc;
      __syncthreads();
      acc = s_acc[0];
#pragma unroll 1
      for( int w = 1; w < nblockthreads / 32; ++w )
         accMerge(acc, s_acc[w]);
      rowTighten<false, SPARSE>(p, acc, sd.x, sd.y, beg, 1, threadIdx.x, nblockthreads, len, s_queue[warp], noalpha);
   }
   if( SPARSE && nnzdone != 0 )
      addRoundNnz(p, nnzdone, gtid >> 5);
}

// ---- short rows that came with their exact activities (flist / facc, see FastAcc): a THREAD per row -------------------
// No activity pass, no butterfly: gates, then one pass over the nonzeros.  The candidate rules are long and few nonzeros
// reach them (the slack test), so the nonzeros that do are COMPACTED over the 32 rows of the warp: they collect in a
// queue in shared memory together with what the rules need to know about their row (FastStash), and the warp runs the
// rules 32 nonzeros at a time with every lane busy.  The rules are the same functions the eight-lanes-per-row path
// calls, on the same numbers (see FastAcc), so the candidates are identical.
constexpr int FAST_CH = 4;            // nonzeros per thread and trip (8 spill: the column words of the next trip must stay in registers)
constexpr int FAST_QCAP = 64;         // fewer than 32 left over + at most 32 new
struct FastStash                      // RowInfo of the row of a lane, as far as candidates() reads it
{
   double lhs, rhs, minact, maxact, maxdelta, slackR, slackL;
   unsigned flags;                    // 1: easy, 2: force, 4: rhs finite, 8: lhs finite
   unsigned spare;
};
struct FastCand
{
   int    rowlane;
   int    colword;
   double a, l, u;
};
struct FastShared
{
   FastStash st[32];
   FastCand  q[FAST_QCAP];
};

__device__ __forceinline__ void fastRunQueue(const Num& n, const Sink& sk, FastShared& sh, int head, int take, int lane, bool& cutoff)
{
   __syncwarp();
   if( lane < take )
   {
      const FastCand c = sh.q[head + lane];
      const FastStash& t = sh.st[c.rowlane];
      RowInfo ri;
      accInit(ri.acc);
      ri.acc.minhi = t.minact;
      ri.acc.maxhi = t.maxact;
      ri.acc.maxdelta = t.maxdelta;
      ri.lhs = t.lhs;
      ri.rhs = t.rhs;
      ri.minact = t.minact;
      ri.maxact = t.maxact;
      ri.slackR = t.slackR;
      ri.slackL = t.slackL;
      ri.easy = (t.flags & 1u) != 0u;
      ri.force = (t.flags & 2u) != 0u;
      ri.rhsfin = (t.flags & 4u) != 0u;
      ri.lhsfin = (t.flags & 8u) != 0u;
      bool touched = false;
      candidates(n, sk, ri, c.a, c.colword & COL_MASK, c.colword < 0, c.l, c.u, cutoff, touched);
      if( touched )
         raiseColumnBit(sk, c.colword & COL_MASK);
   }
   __syncwarp();
}

__device__ __forceinline__ void fastShortPhase(const DevProblem& p, unsigned nfast, int gtid, int nthreads, FastShared& sh)
{
   const Num& n = p.num;
   const int lane = threadIdx.x & 31;
   const unsigned below = (1u << lane) - 1u;
   Sink sk;
   sk.cand = p.cand;
   sk.colbits = p.colbits;
   sk.chglist = p.chglist;
   sk.nchgcols = &p.ctrl->nchgcols;
   sk.listed = false;
   for( unsigned item0 = (unsigned)(gtid - lane); item0 < nfast; item0 += (unsigned)nthreads )      // warp-uniform
   {
      const unsigned item = item0 + (unsigned)lane;
      const bool valid = item < nfast;
      int len = 0;
      long long base = 0;
      bool unit = false;
      bool go = false;
      bool cutoff = false;
      double thr = 0.0;
      RowInfo ri;
      accInit(ri.acc);
      ri.lhs = ri.rhs = 0.0;
      ri.easy = false;
      if( valid )
      {
         const int row = p.flist[item];
         const double2 f0 = reinterpret_cast<const double2*>(p.facc + item)[0];
         const double2 f1 = reinterpret_cast<const double2*>(p.facc + item)[1];
         const double2 sd = reinterpret_cast<const double2*>(p.facc + item)[2];
         len = __double2loint(f1.y);
         base = 32LL * __double2hiint(f1.y) + (row & 31);
         unit = row < p.nsellunit;
         ri.acc.minhi = f0.x;
         ri.acc.maxhi = f0.y;
         ri.acc.maxdelta = f1.x;
         ri.lhs = sd.x;
         ri.rhs = sd.y;
         go = rowGates(n, ri, len, cutoff);
         if( go )
         {
            thr = slackThreshold(n, ri.force);
            FastStash t;
            t.lhs = ri.lhs; t.rhs = ri.rhs; t.minact = ri.acc.minhi; t.maxact = ri.acc.maxhi; t.maxdelta = ri.acc.maxdelta;
            t.slackR = ri.easy ? ri.slackR : 0.0;
            t.slackL = ri.easy ? ri.slackL : 0.0;
            t.flags = (ri.easy ? 1u : 0u) | (ri.force ? 2u : 0u) | (ri.rhsfin ? 4u : 0u) | (ri.lhsfin ? 8u : 0u);
            t.spare = 0u;
            sh.st[lane] = t;
         }
         if( rowInfeasible(n, ri.acc, ri.lhs, ri.rhs) )
            cutoff = true;
      }
      const int golen = go ? len : 0;
      const int kmax = __reduce_max_sync(0xffffffffu, golen);
      int cnt = 0;                       // entries in the queue (warp-uniform)
      // the column words of a trip are fetched while the bounds of the trip before are gathered
      int cjn[FAST_CH];
      // (unconditional loads -- a nonzero that does not exist reads the first one of the row again: with predicated loads the
      // array would live in local memory, and every load would be waited for at once)
#pragma unroll
      for( int q = 0; q < FAST_CH; ++q )
         cjn[q] = p.cols[base + 32LL * (q < golen ? q : 0)];
      for( int k0 = 0; k0 < kmax; k0 += FAST_CH )
      {
         double a[FAST_CH];
         int cj[FAST_CH];
         double2 b[FAST_CH];
#pragma unroll
         for( int q = 0; q < FAST_CH; ++q )
         {
            cj[q] = cjn[q];
            if( k0 + q < golen && !unit )
               a[q] = p.vals[base + 32LL * (k0 + q)];
         }
#pragma unroll
         for( int q = 0; q < FAST_CH; ++q )
            cjn[q] = p.cols[base + 32LL * (k0 + FAST_CH + q < golen ? k0 + FAST_CH + q : 0)];
#pragma unroll
         for( int q = 0; q < FAST_CH; ++q )
         {
            if( k0 + q < golen )
            {
               b[q] = p.bnd[cj[q] & COL_MASK];
               if( unit )
                  a[q] = (cj[q] & COL_NEGCOEF) != 0 ? -1.0 : 1.0;
            }
         }
#pragma unroll
         for( int q = 0; q < FAST_CH; ++q )
         {
            const bool pass = k0 + q < golen && (!ri.easy || passesSlackTest(ri, fabs(a[q]) * (b[q].y - b[q].x), thr));
            const unsigned m = __ballot_sync(0xffffffffu, pass);
            if( m == 0u )
               continue;
            if( pass )
            {
               FastCand c;
               c.rowlane = lane;
               c.colword = cj[q];
               c.a = a[q];
               c.l = b[q].x;
               c.u = b[q].y;
               sh.q[cnt + __popc(m & below)] = c;
            }
            cnt += __popc(m);
            if( cnt >= 32 )
            {
               fastRunQueue(n, sk, sh, 0, 32, lane, cutoff);
               // what is left (fewer than 32 entries) moves to the front
               const int rest = cnt - 32;
               FastCand mv;
               if( lane < rest )
                  mv = sh.q[32 + lane];
               __syncwarp();
               if( lane < rest )
                  sh.q[lane] = mv;
               cnt = rest;
            }
         }
      }
      if( cnt > 0 )
         fastRunQueue(n, sk, sh, 0, cnt, lane, cutoff);
      if( cutoff )
         p.ctrl->cutoff = 1;
      __syncwarp();                      // the stash is rewritten by the next trip
   }
}

// ---- ranged-row propagation (gpulin_ranged.cuh) for the rows marked for propagation, a warp per row ---------------------
// Dense rounds: rangedrow_kernel runs in front of the filter sweeps (which lower the marks); with several GPUs a rank takes
// every nranks-th ranged row.  Small rounds: the marked rows are on the mark lists; rangedListPhase runs in front of the
// exact rules (which claim the rows).
constexpr int RANGED_THREADS = 256;
__global__ void __launch_bounds__(RANGED_THREADS) rangedrow_kernel(const DevProblem p)
{
   const RangedRows& R = p.rr;
   const int gw = (blockIdx.x * RANGED_THREADS + threadIdx.x) >> 5;
   const int nw = (gridDim.x * RANGED_THREADS) >> 5;
   Sink s;
   s.cand = p.cand;
   s.colbits = p.colbits;
   s.chglist = p.chglist;
   s.nchgcols = &p.ctrl->nchgcols;
   s.listed = false;
   unsigned* scratch = R.scratch + (size_t)gw * R.scratchwords;
   for( int i = p.rank + p.nranks * gw; i < R.n; i += p.nranks * nw )
   {
      if( p.dirty[R.row[i]] != ROW_MARKED )
         continue;
      rangedRowWarp(p.num, s, R, p.bnd, p.sides, i, scratch, &p.ctrl->cutoff);
   }
}

// the ranged rows among the rows on the mark lists (all bins); gw / nw: this warp and the number of warps at work
__device__ __forceinline__ void rangedListPhase(const DevProblem& p, const int* ml, unsigned n0, unsigned n1, unsigned n2,
   int gw, int nw)
{
   const RangedRows& R = p.rr;
   if( R.n == 0 )
      return;
   Sink s;
   s.cand = p.cand;
   s.colbits = p.colbits;
   s.chglist = p.chglist;
   s.nchgcols = &p.ctrl->nchgcols;
   s.listed = true;
   unsigned* scratch = R.scratch + (size_t)gw * R.scratchwords;
   const unsigned total = n0 + n1 + n2;
   for( unsigned i = gw; i < total; i += nw )
   {
      const int row = i < n0 ? ml[i] : (i < n0 + n1 ? ml[MARKCAP + (i - n0)] : ml[2 * MARKCAP + (i - n0 - n1)]);
      const int rr = R.idx[row];
      if( rr >= 0 )
         rangedRowWarp(p.num, s, R, p.bnd, p.sides, rr, scratch, &p.ctrl->cutoff);
   }
}

template <int MINB>
__global__ void __launch_bounds__(EXACT_THREADS, MINB) exact_rows_kernel(const DevProblem p)
{
   __shared__ RowAcc s_acc[EXACT_THREADS / 32];
   __shared__ CandQueue s_queue[EXACT_THREADS / 32];

   TRACE_KERNEL_START(p.ctrl, TR_EXACT);
   const unsigned n0 = p.ctrl->nexact[0];
   const unsigned n1 = p.ctrl->nexact[1];
   const unsigned n2 = p.ctrl->nexact[2];
   // (the rows of flist -- the short rows that came with their activities -- are finished by fast_rows_kernel when there are
   // enough of them; otherwise they simply join the short rows of xlist and their activities are computed like everybody
   // else's: a few rows are finished sooner by eight lanes each, in one trip, than by a loop over their nonzeros)
   const unsigned nfast = (p.fastround != 0 && p.ctrl->nexact[3] > p.fastmin) ? 0u : p.ctrl->nexact[3];
   if( (n0 | n1 | n2 | nfast) == 0u )
      return;
   exactPhase<false>(p, p.xlist, n0, p.xlist + p.nsell, n1, p.xlist + p.nsx, n2, s_acc, s_queue, EXACT_THREADS, p.flist, nfast);
}

// runs beside exact_rows_kernel, as its own launch: inside the exact kernel the phase spilled (and slowed the eight-lane path
// down by a tenth).  Two resident blocks per SM: with three or four (80 / 64 registers) the column words of the next trip
// are spilled and waited for at once -- 43 / 48 us instead of 36 for the 74k rows of C3's first round.
constexpr int FAST_THREADS = 256;
__global__ void __launch_bounds__(FAST_THREADS, 2) fast_rows_kernel(const DevProblem p)
{
   __shared__ FastShared s_fast[FAST_THREADS / 32];
   const unsigned nfast = p.ctrl->nexact[3];
   if( nfast <= p.fastmin )
      return;
   if( blockIdx.x == 0 && threadIdx.x == 0 )
      p.ctrl->nfastrows += nfast;
   fastShortPhase(p, nfast, blockIdx.x * FAST_THREADS + threadIdx.x, gridDim.x * FAST_THREADS, s_fast[threadIdx.x >> 5]);
}

// ---- redundancy feedback (propagateCons, cons_linear.c:7743-7753): flags[r] = 1 iff row r (permuted numbering) is
// ---- redundant for the bounds on the device -- exact activities (double-double, inf / huge counters), a warp per row
// ---- whatever its class; not on the propagation path (the plugin asks once per node at most)
__global__ void __launch_bounds__(256) redundant_rows_kernel(const DevProblem p, unsigned char* flags)
{
   const int lane = threadIdx.x & 31;
   const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
   const int nw = (gridDim.x * blockDim.x) >> 5;
   for( int row = gw; row < p.nrows; row += nw )
   {
      const int len = p.rowlen[row] & ROWLEN_MASK;
      const bool sell = row < p.nsell;
      const long long base = sell ? p.sell_off[row >> 5] + (row & 31) : p.rowbeg[row];
      const double2 sd = p.sides[row];
      RowAcc acc;
      accInit(acc);
      double noalpha[ALPHA_SLOTS];
      accumulateExact<false>(p, acc, base, sell ? 32 : 1, lane, 32, len, noalpha);
      accWarpReduce(acc, lane);
      if( lane == 0 )
         flags[row] = rowRedundant(p.num, acc, sd.x, sd.y) ? 1 : 0;
   }
}

// ---- everybody who writes bnd[j] keeps the column's bit in freebits: set iff the bounds are exactly (0,1) ------------
__device__ __forceinline__ bool isFree01(double l, double u)
{
   return l == 0.0 && u == 1.0;
}
__device__ __forceinline__ double2 midHalfWidth(double l, double u)
{
   return make_double2(0.5 * l + 0.5 * u, 0.5 * u - 0.5 * l);
}
__device__ __forceinline__ unsigned char columnState(double l, double u)
{
   if( l == 0.0 )
      return u == 1.0 ? CS_FREE : (u == 0.0 ? CS_FIX0 : CS_OTHER);
   return (l == 1.0 && u == 1.0) ? CS_FIX1 : CS_OTHER;
}
// a bound of a column of integral type that is neither an integer nor infinite switches the FastAcc shortcut off (the flag
// is sticky until the next gpulin_set_bounds; bounds that come out of the propagation itself are integers: adjustedLb/Ub)
__device__ __forceinline__ void checkIntegralBounds(const DevProblem& p, int j, double l, double u)
{
   if( p.coltype[j] != 0 && (l != rint(l) || u != rint(u)) )
      *p.fracflag = 1u;
}
__device__ __forceinline__ void noteBounds(const DevProblem& p, int j, double l, double u)
{
   p.bndf[j] = midHalfWidth(l, u);
   p.colstate[j] = columnState(l, u);
   const unsigned m = 1u << (j & 31);
   if( isFree01(l, u) )
      atomicOr(&p.freebits[j >> 5], m);
   else
      atomicAnd(&p.freebits[j >> 5], ~m);
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
   {
      const_cast<double2*>(p.bnd)[j] = nb;
      noteBounds(p, j, nl, nu);
   }
   return (int)lbchg + (int)ubchg;
}

// marks the rows of column j (entries first, first+step, ... of the calling lane) for the next round
// An entry of colrows carries two flags: a tightened lower / upper bound of the column can matter for a finite side of
// the row (the rule by which eventExecLinear resets boundstightened, cons_linear.c:17239-17252: a larger lower bound
// raises the minimal activity if a > 0 -- that matters if rhs is finite -- or lowers the maximal one if a < 0 -- lhs).
// A row for which the change cannot matter keeps all its residual activities and its verdict: it is not marked.
constexpr int COLROW_LB = 0x40000000;
constexpr int COLROW_UB = (int)0x80000000u;
constexpr int COLROW_ANY = COLROW_LB | COLROW_UB;

// `listfull` (in/out, sticky): a mark list of this round has overflowed -- the next round will be a dense one and nobody
// needs the lists any more.  Then the rows of the thread-per-row class are marked with a blind store (no load in the
// dependent chain); rows of the other classes are still read first: many columns mark the same long rows, and stores to
// one address serialise.
__device__ __forceinline__ bool markListsFull(const DevProblem& p)
{
   const unsigned wb = p.ctrl->mb ^ 1u;
   const uint4 n = __ldcg(reinterpret_cast<const uint4*>(&p.ctrl->nmark[wb][0]));
   return n.x > (unsigned)MARKCAP || n.y > (unsigned)MARKCAP || n.z > (unsigned)MARKCAP;
}

__device__ __forceinline__ void markRowRange(const DevProblem& p, long long qbeg, long long e, int first, int step, int which,
   bool& listfull)
{
   // row ids are fetched four at a time before any flag is stored: the byte stores may alias anything as far as the
   // compiler knows, and a load-store-load-store chain would cost one memory round trip per row
   for( long long q = qbeg + first; q < e; q += 4 * step )
   {
      int r[4];
      long long rb[4];
#pragma unroll
      for( int t = 0; t < 4; ++t )
      {
         r[t] = 0;
         if( q + t * step < e )
            r[t] = p.colrows[q + t * step];
      }
      unsigned char fl[4];
#pragma unroll
      for( int t = 0; t < 4; ++t )
      {
         fl[t] = ROW_MARKED;
         const bool matters = (r[t] & which) != 0;
         r[t] &= ~COLROW_ANY;
         if( q + t * step < e && matters )
         {
            if( listfull && r[t] < p.nsell )
               fl[t] = ROW_CLEAN;          // blind store below
            else
            {
               fl[t] = p.dirty[r[t]];
               if( r[t] >= p.nsell && r[t] < p.nsx )
                  rb[t] = p.rowbeg[r[t]];
            }
         }
      }
      // the flag goes up with an atomic on its word: exactly one of the columns that race for a row notes it, so the
      // lists hold no row twice and their lengths are the same on every rank of a node (the ranks decide by them
      // whether the next round is a small one, and must decide alike); the atomics of a trip are in flight together
      bool won[4];
#pragma unroll
      for( int t = 0; t < 4; ++t )
      {
         won[t] = false;
         if( fl[t] != ROW_MARKED )
         {
            if( listfull && r[t] < p.nsell )
               p.dirty[r[t]] = ROW_MARKED;            // (nobody reads the lists of this round any more)
            else
            {
               unsigned* word = reinterpret_cast<unsigned*>(p.dirty + (r[t] & ~3));
               const int shift = 8 * (r[t] & 3);
               won[t] = ((atomicOr(word, (unsigned)ROW_MARKED << shift) >> shift) & 0xffu) != ROW_MARKED;
            }
         }
      }
#pragma unroll
      for( int t = 0; t < 4; ++t )
      {
         if( won[t] )
         {
            if( r[t] >= p.nsell && r[t] < p.nsx )
               p.tileflag[(rb[t] - p.streambase) >> 8] = 1;
            if( listfull )
               continue;
            const int bin = r[t] < p.nsell ? 0 : (r[t] < p.nsx ? 1 : 2);
            const unsigned wb = p.ctrl->mb ^ 1u;
            const cooperative_groups::coalesced_group g = cooperative_groups::coalesced_threads();
            const unsigned same = g.match_any(bin);
            const int leader = __ffs(same) - 1;
            unsigned pos = 0u;
            if( (int)g.thread_rank() == leader )
               pos = atomicAdd(&p.ctrl->nmark[wb][bin], (unsigned)__popc(same));
            pos = g.shfl(pos, leader) + (unsigned)__popc(same & ((1u << g.thread_rank()) - 1u));
            if( pos < (unsigned)MARKCAP )
               p.marklist[(wb * 3 + bin) * MARKCAP + pos] = r[t];
            else
               listfull = true;
         }
      }
   }
}

__device__ __forceinline__ void markColumnRows(const DevProblem& p, int j, int first, int step, int which, bool& listfull)
{
   markRowRange(p, p.colbeg[j], p.colbeg[j + 1], first, step, which, listfull);
}

__device__ __forceinline__ void markColumnRows(const DevProblem& p, int j, int first, int step, int which = COLROW_ANY)
{
   bool listfull = false;
   markColumnRows(p, j, first, step, which, listfull);
}

// loop control (propagateDomains, solve.c:766-787), run by ONE WARP after the last column was applied (all 32 lanes call
// it): the lanes add up the spread counters of swept nonzeros, lane 0 does the rest -- every load is issued before the
// first store, so that the step costs two memory round trips instead of one per field
template <bool GRAPH>
__device__ __forceinline__ void controlStep(Ctrl* c, cudaGraphConditionalHandle handle)
{
   static_assert(NNZ_SLOTS == 64, "two slots per lane");
   const int lane = threadIdx.x & 31;
   unsigned long long nnz = __ldcg(&c->round_nnz[lane]) + __ldcg(&c->round_nnz[lane + 32]);
   const unsigned long long nchg = __ldcg(&c->round_nchg);
   const int r = __ldcg(&c->round);
   const int cutoff = __ldcg(&c->cutoff);
   const int maxrounds = __ldcg(&c->maxrounds);
   const unsigned mb = __ldcg(&c->mb);
   const unsigned long long tnchg = __ldcg(&c->total_nchg);
   const unsigned long long tnnz = __ldcg(&c->total_nnz);
   c->round_nnz[lane] = 0;
   c->round_nnz[lane + 32] = 0;
#pragma unroll
   for( int m = 16; m >= 1; m >>= 1 )
      nnz += __shfl_xor_sync(0xffffffffu, nnz, m);
   if( lane != 0 )
      return;
   if( r < MAX_HIST )
   {
      c->hist_time[r] = globaltimer();
      c->hist_nnz[r] = nnz;
      c->hist_nchg[r] = nchg;
      if( r + 1 < MAX_HIST )
         c->hist_push[r + 1] = 0;       // (set by the exchange of that round, if it has one)
   }
   c->total_nchg = tnchg + nchg;
   c->total_nnz = tnnz + nnz;
   c->round_nchg = 0;
   c->ticket = 0;
   c->nchgcols = 0;
   c->nexact[0] = c->nexact[1] = c->nexact[2] = c->nexact[3] = 0;
   // the list this apply step filled becomes the one a sparse round reads; the other one is empty again
   c->nmark[mb][0] = c->nmark[mb][1] = c->nmark[mb][2] = 0;
   c->mb = mb ^ 1u;
   c->round = r + 1;
   int cont = 0;
   if( cutoff )
      c->status = 1;
   else if( nchg == 0 )
      c->status = 0;
   else if( maxrounds > 0 && r + 1 >= maxrounds )
      c->status = 2;
   else
      cont = 1;
   c->cont = cont;
   if( GRAPH )
      cudaGraphSetConditional(handle, (unsigned)cont);
}

constexpr int APPLY_THREADS = 256;
constexpr int APPLY_G = 4;           // lanes per changed column in the apply kernel (rounds with many changes)
constexpr int SPARSE_G = 8;          // ... in the sparse-rounds kernel (few changes: latency counts)

// appends the accepted changes of column j to the round-ordered change log
__device__ __forceinline__ void logChanges(const DevProblem& p, int j, int round, int logcap, int nc, bool lbchg, bool ubchg,
   const double2& nb)
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

// the same for a whole warp (all 32 lanes call it): one atomic on the log counter per warp instead of one per column --
// a round with 250k changed columns spent 0.19 ms on that one address
__device__ __forceinline__ void logChangesWarp(const DevProblem& p, int j, int round, int logcap, bool lbchg, bool ubchg,
   const double2& nb)
{
   const unsigned lane = threadIdx.x & 31u;
   const unsigned below = (1u << lane) - 1u;
   const unsigned mlb = __ballot_sync(0xffffffffu, lbchg);
   const unsigned mub = __ballot_sync(0xffffffffu, ubchg);
   const int total = __popc(mlb) + __popc(mub);
   if( total == 0 )
      return;
   unsigned long long pos = 0;
   if( lane == 0 )
      pos = atomicAdd(&p.ctrl->logcount, (unsigned long long)total);
   pos = __shfl_sync(0xffffffffu, pos, 0) + (unsigned long long)(__popc(mlb & below) + __popc(mub & below));
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

// Buffered variant for rounds with many changes (apply_kernel): the records of a warp collect in its shared-memory buffer
// over the trips and go to the log in one piece -- one atomic per LOGBUF records or so, and coalesced stores.
constexpr int LOGBUF = 96;           // records per warp; a trip adds at most 2 * 32 / APPLY_G = 16
struct WarpLog
{
   ChangeRec* buf;                   // LOGBUF records of this warp (NULL: unbuffered, logChangesWarp)
   int        n;                     // records in the buffer (warp-uniform)
};

__device__ __forceinline__ void flushWarpLog(const DevProblem& p, WarpLog& w, int logcap)
{
   if( w.n == 0 )
      return;
   const unsigned lane = threadIdx.x & 31u;
   unsigned long long base = 0;
   if( lane == 0 )
      base = atomicAdd(&p.ctrl->logcount, (unsigned long long)w.n);
   base = __shfl_sync(0xffffffffu, base, 0);
   __syncwarp();
   for( int i = (int)lane; i < w.n; i += 32 )
   {
      if( base + (unsigned long long)i < (unsigned long long)logcap )
         p.log[base + (unsigned long long)i] = w.buf[i];
   }
   __syncwarp();
   w.n = 0;
}

__device__ __forceinline__ void logChangesBuffered(const DevProblem& p, WarpLog& w, int j, int round, int logcap, bool lbchg,
   bool ubchg, const double2& nb)
{
   const unsigned lane = threadIdx.x & 31u;
   const unsigned below = (1u << lane) - 1u;
   const unsigned mlb = __ballot_sync(0xffffffffu, lbchg);
   const unsigned mub = __ballot_sync(0xffffffffu, ubchg);
   const int total = __popc(mlb) + __popc(mub);
   if( total == 0 )
      return;
   if( w.n + total > LOGBUF )
      flushWarpLog(p, w, logcap);
   int pos = w.n + __popc(mlb & below) + __popc(mub & below);
   if( lbchg )
   {
      ChangeRec rec;
      rec.var = j; rec.round = round; rec.newbound = nb.x; rec.is_upper = 0; rec.reserved = 0;
      w.buf[pos++] = rec;
   }
   if( ubchg )
   {
      ChangeRec rec;
      rec.var = j; rec.round = round; rec.newbound = nb.y; rec.is_upper = 1; rec.reserved = 0;
      w.buf[pos] = rec;
   }
   w.n += total;
}

// the columns on the change list of this round, eight lanes per column: one accepts the bounds, all mark the rows of the
// column; returns the number of bound changes this thread accepted
template <int G>
__device__ __forceinline__ int applyListPhase(const DevProblem& p, unsigned nlist, int gtid, int nthreads, int round, int logcap,
   ChangeRec* warplogbuf = nullptr)
{
   WarpLog wlog;
   wlog.buf = warplogbuf;
   wlog.n = 0;
   const int gl = gtid & (G - 1);
   const int ngroups = nthreads / G;
   const unsigned trips = (nlist + ngroups - 1) / ngroups;       // warp-uniform
   int mychg = 0;
   bool listfull = false;
   // software pipeline over the trips of a group: the column index is fetched two trips ahead, its keys, bounds and row
   // range one trip ahead -- the apply step is nothing but dependent loads (chglist -> cand / bnd / colbeg -> colrows ->
   // flags), and with a handful of trips per group their latencies add up
   const unsigned item0 = gtid / G;
   int jB = item0 < nlist ? p.chglist[item0] : -1;                       // column of trip `it`
   int jA = item0 + ngroups < nlist ? p.chglist[item0 + ngroups] : -1;   // column of trip `it + 1`
   longlong2 kB = make_longlong2(0, 0);
   double2 oldB = make_double2(0.0, 0.0);
   long long cb0 = 0;
   long long cb1 = 0;
   if( jB >= 0 )
   {
      kB = reinterpret_cast<const longlong2*>(p.cand)[jB];
      oldB = p.bnd[jB];
      cb0 = p.colbeg[jB];
      cb1 = p.colbeg[jB + 1];
   }
   for( unsigned it = 0; it < trips; ++it )
   {
      const int j = jB;
      const bool valid = j >= 0;
      const longlong2 k = kB;
      const double2 old = oldB;
      const long long q0 = cb0;
      const long long q1 = cb1;
      // next trips
      jB = jA;
      {
         const unsigned item2 = (it + 2) * ngroups + item0;
         jA = (it + 2 < trips && item2 < nlist) ? p.chglist[item2] : -1;
      }
      if( jB >= 0 )
      {
         kB = reinterpret_cast<const longlong2*>(p.cand)[jB];
         oldB = p.bnd[jB];
         cb0 = p.colbeg[jB];
         cb1 = p.colbeg[jB + 1];
      }
      if( !listfull )
         listfull = markListsFull(p);        // one load per trip, beside the loads above (not once per marked row)
      // which of the two bounds moves says which rows can care (every lane of the group holds the two words itself,
      // read BEFORE the first lane accepts the bounds)
      const int which = valid ? ((key2d(~k.x) != old.x ? COLROW_LB : 0) | (key2d(k.y) != old.y ? COLROW_UB : 0)) : 0;
      bool lbchg = false;
      bool ubchg = false;
      double2 nb = make_double2(0.0, 0.0);
      if( valid && gl == 0 )
      {
         const int nc = applyColumn(p, j, nb, lbchg, ubchg);
         atomicAnd(&p.colbits[j >> 5], ~(1u << (j & 31)));
         mychg += nc;
      }
      if( logcap > 0 )
      {
         if( wlog.buf != nullptr )
            logChangesBuffered(p, wlog, j, round, logcap, lbchg, ubchg, nb);
         else
            logChangesWarp(p, j, round, logcap, lbchg, ubchg, nb);
      }
      // every candidate that reached the column beat the round-start bound, so the column changes (a crossing pair
      // clamped back to its old value is the one exception): the group marks without waiting for the verdict
      if( valid )
         markRowRange(p, q0, q1, gl, G, which, listfull);
   }
   if( wlog.buf != nullptr )
      flushWarpLog(p, wlog, logcap);
   return mychg;
}

// DENSE = false: the columns on the change list of this round -- cost proportional to the changes (single GPU);
//                eight lanes per column: one accepts the bounds, all mark the rows of the column;
// DENSE = true : every column compares its (all-reduced) candidate keys with its bounds (rows sharded over ranks:
//                a key may have been moved by another rank: the host-driven NCCL rounds)
constexpr int APPLY_LIST = 0;
constexpr int APPLY_DENSE = 1;

template <int MODE, bool GRAPH>
__global__ void __launch_bounds__(APPLY_THREADS, 4) apply_kernel(const DevProblem p, cudaGraphConditionalHandle handle)
{
   __shared__ int s_nchg;
   __shared__ ChangeRec s_log[MODE == APPLY_LIST ? APPLY_THREADS / 32 : 1][MODE == APPLY_LIST ? LOGBUF : 1];
   if( threadIdx.x == 0 )
      s_nchg = 0;
   __syncthreads();

   Ctrl* c = p.ctrl;
   TRACE_KERNEL_START(c, TR_APPLY);
   const int lane = threadIdx.x & 31;
   const int round = c->round;
   const int logcap = c->logcap;
   const unsigned nlist = c->nchgcols;
   int mychg = 0;
   const int nthreads = gridDim.x * APPLY_THREADS;
   const int gtid = blockIdx.x * APPLY_THREADS + threadIdx.x;
   if( MODE == APPLY_DENSE )
   {
      // the local list only serves to lower the bits again
      for( unsigned i = gtid; i < nlist; i += nthreads )
      {
         const int j = p.chglist[i];
         atomicAnd(&p.colbits[j >> 5], ~(1u << (j & 31)));
      }
      for( int j = gtid; j < p.ncols; j += nthreads )
      {
         bool lbchg;
         bool ubchg;
         double2 nb;
         const int nc = applyColumn(p, j, nb, lbchg, ubchg);
         if( nc > 0 )
         {
            markColumnRows(p, j, 0, 1, (lbchg ? COLROW_LB : 0) | (ubchg ? COLROW_UB : 0));
            if( logcap > 0 )
               logChanges(p, j, round, logcap, nc, lbchg, ubchg, nb);
         }
         mychg += nc;
      }
   }
   else
   {
      // Many changed columns (one mark per row and more to expect): walking the rows of every changed column costs more than
      // the filter sweep of the rows that would have stayed unmarked -- ALL rows are marked instead (a row whose bounds did
      // not move is finished by the filter, or reproduces candidates that are in place already: the result is the same).
      // The decision depends on the length of the change list alone, so every rank of a node takes it alike.
      const bool markall = nlist >= p.markall_min;
      if( !markall )
         mychg += applyListPhase<APPLY_G>(p, nlist, gtid, nthreads, round, logcap, s_log[MODE == APPLY_LIST ? threadIdx.x >> 5 : 0]);
      else
      {
         // a thread per column: accept the bounds, log the changes
         WarpLog wlog;
         wlog.buf = s_log[MODE == APPLY_LIST ? threadIdx.x >> 5 : 0];
         wlog.n = 0;
         const unsigned trips = (nlist + nthreads - 1) / nthreads;       // uniform
         for( unsigned it = 0; it < trips; ++it )
         {
            const unsigned i = it * nthreads + gtid;
            const int j = i < nlist ? p.chglist[i] : -1;
            bool lbchg = false;
            bool ubchg = false;
            double2 nb = make_double2(0.0, 0.0);
            if( j >= 0 )
            {
               mychg += applyColumn(p, j, nb, lbchg, ubchg);
               atomicAnd(&p.colbits[j >> 5], ~(1u << (j & 31)));
            }
            if( logcap > 0 )
               logChangesBuffered(p, wlog, j, round, logcap, lbchg, ubchg, nb);
         }
         flushWarpLog(p, wlog, logcap);
      }
      if( markall )
      {
         uint4* d = reinterpret_cast<uint4*>(p.dirty);
         const unsigned one = 0x01010101u * ROW_MARKED;
         const int nq = p.nrows >> 4;
         for( int i = gtid; i < nq; i += nthreads )
            d[i] = make_uint4(one, one, one, one);
         for( int r = (nq << 4) + gtid; r < p.nrows; r += nthreads )
            p.dirty[r] = ROW_MARKED;
         for( int t = gtid; t < p.ntiles; t += nthreads )
            p.tileflag[t] = 1;
         if( gtid == 0 )
            atomicMax(&c->nmark[c->mb ^ 1u][0], (unsigned)MARKCAP + 1u);      // the lists are not complete: a dense round follows
      }
   }
   mychg = __reduce_add_sync(0xffffffffu, mychg);
   if( lane == 0 && mychg != 0 )
      atomicAdd(&s_nchg, mychg);
   __syncthreads();
   __shared__ bool s_last;
   if( threadIdx.x == 0 )
   {
      if( s_nchg != 0 )
         atomicAdd(&c->round_nchg, (unsigned long long)s_nchg);
      __threadfence();
      s_last = atomicAdd(&c->ticket, 1u) == gridDim.x - 1;
   }
   __syncthreads();
   if( s_last && threadIdx.x < 32 )
   {
      __threadfence();
      controlStep<GRAPH>(c, handle);
   }
}

// ---- dense rounds: the changed-column bits become the change list -- and travel to the other GPUs of the node -----------
// In a dense round the exact kernel commits with reductions only (atomicMin on the key, atomicOr on the column's bit:
// nobody waits for an answer).  collect_kernel turns the bits into the change list of the apply step: a word per thread,
// one atomic on the list counter per warp.
//
// PEERS -- dense rounds sharded over the GPUs of a node.  Every rank holds the WHOLE matrix (0.25 - 0.9 GB of 180) and
// identical bounds, sweeps only its share of the rows (DevProblem::rank ... st1) and commits the candidates of those rows
// into its own key vector.  Then
//   collect_kernel<true>  stores every column it lists, (column, ~key(lb), key(ub)), into the inbox of every other rank
//                         through peer memory (NVLink; plain stores that validate themselves, see PeerEntry: no fence,
//                         no atomics), the last block to finish adds the header (count, verdict);
//   peer_merge_kernel     waits for the headers of all other ranks, merges their entries into the local keys with
//                         atomicMin and puts columns on the change list that are not there yet.
// (Measured on 8 GPUs, C4's first round, 85k entries per rank: bulk copies 137 us, plain 16-byte stores in 512-byte runs
// 278 us; letting the receivers PULL the lists from the senders' memory instead took 20 us to list and 300 us to merge.)
// After the merge all ranks hold the same keys and the same set of changed columns; the list-driven apply step and the
// small rounds that follow (sparse_rounds_kernel) run on every rank redundantly -- deterministic, so the ranks stay
// identical and need no further exchange until the next dense round.  One exchange per dense round, volume proportional
// to the changed columns; its counterpart in the reference is the min/max merge of syncstore.c:921.
// The inboxes are double buffered by the parity of the exchange number: a rank can be at most one exchange ahead of
// the slowest rank (it needs everybody's entries of exchange e before it can send those of e + 1).
__device__ __forceinline__ void storeV4(void* addr, const uint4& v)
{
   asm volatile("st.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 loadV4Volatile(const void* addr)
{
   uint4 v;
   asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(addr) : "memory");
   return v;
}

// shared -> global bulk copy through the async proxy (TMA); the destination may be peer memory: the copy engine of the SM
// writes whole bursts to the link instead of one packet per warp store
__device__ __forceinline__ void bulkStore(void* dst, unsigned src, unsigned bytes)
{
   asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
constexpr int COLLECT_WARPS = 256 / 32;
constexpr int COLLECT_CHUNK = 128;                      // entries a warp stages before it sends them (4 KB)
constexpr int COLLECT_SMEM_PEERS = COLLECT_WARPS * (1024 * (int)sizeof(int) + COLLECT_CHUNK * (int)sizeof(PeerEntry));

template <bool PEERS>
__global__ void __launch_bounds__(256) collect_kernel(const DevProblem p)
{
   Ctrl* c = p.ctrl;
   TRACE_KERNEL_START(c, TR_PUSH);
   const int lane = threadIdx.x & 31;
   const int nwords = (p.ncols + 31) >> 5;
   const int nthreads = gridDim.x * blockDim.x;
   const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
   const unsigned epoch = c->epoch + 1u;          // (PEERS: the last block to finish stores it)
   const int parity = (int)(epoch & 1u);
   if( PEERS && gtid == 0 && c->round < MAX_HIST )
   {
      c->hist_push[c->round] = globaltimer();
      c->hist_wait[c->round] = 0;
   }
   // per warp: the columns of its 32 words in list order (and, with peers, the entries of a chunk on their way out)
   __shared__ int s_stage[PEERS ? 1 : COLLECT_WARPS][PEERS ? 1 : 1024];
   extern __shared__ __align__(128) unsigned char s_dyn[];
   int* stage = PEERS ? reinterpret_cast<int*>(s_dyn) + (threadIdx.x >> 5) * 1024 : s_stage[PEERS ? 0 : (threadIdx.x >> 5)];
   PeerEntry* sent = reinterpret_cast<PeerEntry*>(s_dyn + COLLECT_WARPS * 1024 * sizeof(int)) + (threadIdx.x >> 5) * COLLECT_CHUNK;
   for( int w0 = gtid - lane; w0 < nwords; w0 += nthreads )      // warp-uniform trip count
   {
      const int w = w0 + lane;
      unsigned bits = w < nwords ? p.colbits[w] : 0u;
      const int mine = __popc(bits);
      int incl = mine;
#pragma unroll
      for( int d = 1; d < 32; d <<= 1 )
      {
         const int o = __shfl_up_sync(0xffffffffu, incl, d);
         if( lane >= d )
            incl += o;
      }
      const int total = __shfl_sync(0xffffffffu, incl, 31);
      if( total == 0 )
         continue;
      unsigned base = 0u;
      if( lane == 31 )
         base = atomicAdd(&c->nchgcols, (unsigned)total);
      base = __shfl_sync(0xffffffffu, base, 31);
      // the columns are staged in list order, so that what follows -- the stores to the list, the entries for the peers --
      // is contiguous over the lanes of the warp
      int pos = incl - mine;
      while( bits != 0u )
      {
         stage[pos++] = 32 * w + __ffs(bits) - 1;
         bits &= bits - 1u;
      }
      __syncwarp();
      if( !PEERS )
      {
         for( int i = lane; i < total; i += 32 )
            p.chglist[base + i] = stage[i];
      }
      else
      {
         const PeerTable& t = *p.peers;
         for( int c0 = 0; c0 < total; c0 += COLLECT_CHUNK )      // uniform
         {
            const int n = min(COLLECT_CHUNK, total - c0);
            int j[COLLECT_CHUNK / 32];
            longlong2 k[COLLECT_CHUNK / 32];
#pragma unroll
            for( int q = 0; q < COLLECT_CHUNK / 32; ++q )
            {
               const int i = c0 + 32 * q + lane;
               j[q] = i < total ? stage[i] : -1;
               if( j[q] >= 0 )
               {
                  p.chglist[base + i] = j[q];
                  k[q] = reinterpret_cast<const longlong2*>(p.cand)[j[q]];
               }
            }
#pragma unroll
            for( int q = 0; q < COLLECT_CHUNK / 32; ++q )
            {
               if( j[q] >= 0 )
               {
                  PeerEntry e;
                  e.a = make_uint4((unsigned)j[q], epoch, (unsigned)(unsigned long long)k[q].x, (unsigned)((unsigned long long)k[q].x >> 32));
                  e.b = make_uint4((unsigned)(unsigned long long)k[q].y, (unsigned)((unsigned long long)k[q].y >> 32), epoch, 0u);
                  sent[32 * q + lane] = e;
               }
            }
            // the entries of the chunk leave as one bulk copy per peer (generic-proxy writes to shared memory first become
            // visible to the async proxy); every rank starts with its right neighbour: eight ranks that all write to rank 0
            // first, then to rank 1, ... would meet at one port of the switch
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if( lane == 0 )
            {
               for( int kk = 1; kk < t.n; ++kk )
               {
                  const int r = (t.rank + kk) % t.n;
                  bulkStore(peerEntries(t, t.box[r], parity, t.rank) + base + c0, smemAddr(sent), (unsigned)n * (unsigned)sizeof(PeerEntry));
               }
               asm volatile("cp.async.bulk.commit_group;" ::: "memory");
               asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");      // the chunk buffer may be refilled
            }
            __syncwarp();
         }
      }
      __syncwarp();
   }
   if( PEERS && lane == 0 )
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
   if( !PEERS )
      return;
   // the marks of this round are spent on every rank: the rows of the other ranks' shares were swept there
   {
      uint4* d = reinterpret_cast<uint4*>(p.dirty);
      const int nq = (p.nrows + 15) >> 4;
      for( int i = gtid; i < nq; i += nthreads )
         d[i] = make_uint4(0u, 0u, 0u, 0u);
      uint4* f = reinterpret_cast<uint4*>(p.tileflag);
      const int nf = (p.ntiles + 15) >> 4;
      for( int i = gtid; i < nf; i += nthreads )
         f[i] = make_uint4(0u, 0u, 0u, 0u);
   }
   __threadfence();
   __syncthreads();
   __shared__ bool s_last;
   if( threadIdx.x == 0 )
      s_last = atomicAdd(&c->pushticket, 1u) == gridDim.x - 1;
   __syncthreads();
   if( !s_last )
      return;
   __threadfence();
   const PeerTable& t = *p.peers;
   const unsigned n = *reinterpret_cast<volatile unsigned*>(&c->nchgcols);
   if( (int)threadIdx.x < t.n && (int)threadIdx.x != t.rank )
   {
      unsigned long long nnz = 0;
      for( int i = 0; i < NNZ_SLOTS; ++i )
         nnz += c->round_nnz[i];
      PeerHeader* h = peerHeader(t.box[threadIdx.x], parity, t.rank);
      *reinterpret_cast<volatile unsigned long long*>(&h->nnz) = nnz;
      *reinterpret_cast<volatile unsigned long long*>(&h->word) =
         ((unsigned long long)epoch << 32) | (c->cutoff ? 0x80000000ull : 0ull) | (unsigned long long)n;
   }
   __syncthreads();
   if( threadIdx.x == 0 )
   {
      c->pushticket = 0;
      c->epoch = epoch;
      traceStamp(c, TR_PUSH_END);
   }
}

__global__ void __launch_bounds__(256) peer_merge_kernel(const DevProblem p)
{
   const PeerTable& t = *p.peers;
   Ctrl* c = p.ctrl;
   const unsigned epoch = c->epoch;
   const int parity = (int)(epoch & 1u);
   __shared__ unsigned s_count[MAX_PEERS];
   __shared__ int s_fail;
   TRACE_KERNEL_START(c, TR_MERGE);
   if( threadIdx.x == 0 )
      s_fail = 0;
   __syncthreads();
   if( (int)threadIdx.x < t.n )
   {
      unsigned cnt = 0u;
      if( (int)threadIdx.x != t.rank )
      {
         PeerHeader* h = peerHeader(t.box[t.rank], parity, threadIdx.x);
         const unsigned long long tstart = globaltimer();
         unsigned long long word;
         for( ;; )
         {
            word = *reinterpret_cast<volatile unsigned long long*>(&h->word);
            if( (unsigned)(word >> 32) == epoch )
               break;
            if( globaltimer() - tstart > 5000000000ull )     // 5 s: a peer is gone -- give up instead of hanging the GPU
            {
               s_fail = 1;
               break;
            }
         }
         if( blockIdx.x == 0 && c->round < MAX_HIST )
            atomicMax(&c->hist_wait[c->round], globaltimer() - tstart);
         if( !s_fail )
         {
            cnt = (unsigned)word & 0x7fffffffu;
            if( (word & 0x80000000ull) != 0ull )
               c->cutoff = 1;
            if( blockIdx.x == 0 )
               addRoundNnz(p, *reinterpret_cast<volatile unsigned long long*>(&h->nnz), threadIdx.x);
         }
      }
      s_count[threadIdx.x] = cnt;
   }
   __syncthreads();
   TRACE_KERNEL_START(c, TR_MERGE_READY);
   Sink s;
   s.cand = p.cand;
   s.colbits = p.colbits;
   s.chglist = p.chglist;
   s.nchgcols = &p.ctrl->nchgcols;
   s.listed = true;
   // the lists of all sources as one index space (entry i belongs to the source r with off[r] <= i < off[r + 1]): every
   // thread of the grid has an entry in flight, whatever the number of sources and the lengths of their lists
   unsigned off[MAX_PEERS + 1];
   off[0] = 0u;
#pragma unroll
   for( int r = 0; r < MAX_PEERS; ++r )
      off[r + 1] = off[r] + (r < t.n ? s_count[r] : 0u);
   const unsigned total = off[MAX_PEERS];
   const unsigned nthreads = gridDim.x * blockDim.x;
   for( unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total && !s_fail; i += nthreads )
   {
      int r = 0;
      unsigned o = 0u;
#pragma unroll
      for( int q = 1; q < MAX_PEERS; ++q )
      {
         if( i >= off[q] )
         {
            r = q;
            o = off[q];
         }
      }
      const PeerEntry* ent = peerEntries(t, t.box[t.rank], parity, r) + (i - o);
      // the entry may still be on its way: both halves carry the exchange number once they have landed
      uint4 a = loadV4Volatile(&ent->a);
      uint4 b = loadV4Volatile(&ent->b);
      if( a.y != epoch || b.z != epoch )
      {
         const unsigned long long tstart = globaltimer();
         do
         {
            a = loadV4Volatile(&ent->a);
            b = loadV4Volatile(&ent->b);
         } while( (a.y != epoch || b.z != epoch) && globaltimer() - tstart < 5000000000ull );
         if( a.y != epoch || b.z != epoch )
         {
            s_fail = 1;
            break;
         }
      }
      const int j = (int)a.x;
      const long long kl = (long long)(((unsigned long long)a.w << 32) | a.z);
      const long long ku = (long long)(((unsigned long long)b.y << 32) | b.x);
      atomicMin(&p.cand[2 * (size_t)j], kl);
      atomicMin(&p.cand[2 * (size_t)j + 1], ku);
      const bool first = raiseColumnBit(s, j);
      listChangedColumn(s, j, first);
   }
   __syncthreads();
   if( s_fail && threadIdx.x == 0 )
   {
      c->peererror = 1;
      c->cutoff = 1;
   }
}

// ---- sparse rounds: a persistent cooperative kernel ----------------------------------------------------------------
// Once the bounds have nearly settled a round touches a few hundred rows, and three launches plus a graph-loop iteration
// cost more than the work.  This kernel follows the apply step inside the loop body: as long as the last apply marked
// at most SPARSE_MAXROWS rows (all of them on the mark list) it runs whole rounds by itself -- exact rules for the marked
// rows, grid sync, apply for the changed columns, grid sync, loop control -- and leaves when a round is not sparse any
// more or the loop ends.  The filter pass is skipped in these rounds: every marked row gets the exact rules.
constexpr int SPARSE_THREADS = 256;
constexpr unsigned SPARSE_MAXROWS = 1u << 14;      // short rows (8 lanes each)
constexpr unsigned SPARSE_MAXMEDIUM = 1u << 12;    // rows of 33..4096 nonzeros (a warp each)

template <bool GRAPH>
__global__ void __launch_bounds__(SPARSE_THREADS) sparse_rounds_kernel(const DevProblem p, cudaGraphConditionalHandle handle)
{
   __shared__ RowAcc s_acc[SPARSE_THREADS / 32];
   __shared__ CandQueue s_queue[SPARSE_THREADS / 32];
   __shared__ int s_nchg;
   cooperative_groups::grid_group grid = cooperative_groups::this_grid();
   Ctrl* c = p.ctrl;
   const int lane = threadIdx.x & 31;
   const int gtid = blockIdx.x * SPARSE_THREADS + threadIdx.x;
   const int nthreads = gridDim.x * SPARSE_THREADS;
   TRACE_KERNEL_START(c, TR_SPARSE);

   for( ;; )
   {
      // every thread reads the same state: it was written before the previous grid sync (or by the apply kernel)
      const unsigned mb = c->mb;
      const unsigned n0 = c->nmark[mb][0];
      const unsigned n1 = c->nmark[mb][1];
      const unsigned n2 = c->nmark[mb][2];
      const bool sparse = c->cont != 0 && n0 <= (unsigned)MARKCAP && n1 <= (unsigned)MARKCAP && n2 <= (unsigned)MARKCAP
         && n0 <= SPARSE_MAXROWS && n1 <= SPARSE_MAXMEDIUM && n2 <= 64u;
      if( !sparse )
         break;
      if( threadIdx.x == 0 )
         s_nchg = 0;

      // ---- the exact rules for the marked rows (ranged rows: the gcd rule first, it does not claim the row)
      const int* ml = p.marklist + (size_t)mb * 3 * MARKCAP;
      rangedListPhase(p, ml, n0, n1, n2, gtid >> 5, nthreads >> 5);
      exactPhase<true>(p, ml, n0, ml + MARKCAP, n1, ml + 2 * MARKCAP, n2, s_acc, s_queue, SPARSE_THREADS);
      __threadfence();
      grid.sync();

      // ---- accept the changes, mark the rows of the changed columns (into the other mark list)
      const int round = c->round;
      const unsigned nlist = c->nchgcols;
      int mychg = applyListPhase<SPARSE_G>(p, nlist, gtid, nthreads, round, c->logcap);
      mychg = __reduce_add_sync(0xffffffffu, mychg);
      if( lane == 0 && mychg != 0 )
         atomicAdd(&s_nchg, mychg);
      __syncthreads();
      // ---- loop control: by the block that finishes the apply phase last, in front of the second grid sync
      __shared__ bool s_last;
      if( threadIdx.x == 0 )
      {
         if( s_nchg != 0 )
            atomicAdd(&c->round_nchg, (unsigned long long)s_nchg);
         __threadfence();
         s_last = atomicAdd(&c->ticket, 1u) == gridDim.x - 1;
      }
      __syncthreads();
      if( s_last && threadIdx.x < 32 )
      {
         __threadfence();
         controlStep<GRAPH>(c, handle);
         if( threadIdx.x == 0 )
            ++c->nsparse;
         __threadfence();
      }
      grid.sync();
   }
}

// ---- one probe = one launch of one block ----------------------------------------------------------------------------
// SCIPbacktrackProbing + SCIPchgVarLb/UbProbing + SCIPpropagateProbing (scip_probing.c:226/:302/:346/:581) for a worker
// clone: undo the change log of the previous probe of this worker, set the probed bound, mark the rows of the column and
// run sparse rounds (exact rules for the marked rows, apply, loop control) inside the block until the fixpoint, a
// cutoff or the round limit -- no other launch, no host round trip; the verdict goes to out[0].  A probe is a small
// cascade (a few dozen rows); one that outgrows the block (more than PROBE_MAXROWS marked rows in a round, a long row,
// a full change log) stops with GPULIN_PROBE_OVERFLOW and marks the worker as poisoned: the host reruns such probes
// through the general loop after a reset.  Whatever is still marked when the probe ends is unmarked again, so that the
// next probe of this worker starts from "node + change log".
constexpr int PROBE_THREADS = 256;
constexpr unsigned PROBE_MAXROWS = 4096;
constexpr unsigned PROBE_MAXMEDIUM = 64;
constexpr int GPULIN_PROBE_OVERFLOW = 3;

struct ProbeResult
{
   int       status;
   int       nrounds;
   long long nchanges;
   long long logoff;     // position of the probe's change log in the output of the batch (-1: it did not fit, or none was asked for)
   long long nlog;       // its length
};

// The same kernel serves small incremental calls on any handle (j < 0, keepmarks): the rows gpulin_update_bounds marked
// since the last fixpoint are on mark list mb^1; if the cascade outgrows the block the marks stay and the general loop
// continues the call (Ctrl::resume).
__device__ __forceinline__ void probeBody(const DevProblem& p, const DevProblem& base, int restorevar, int j, double l, double u,
   int maxrounds, int logcap, int keepmarks, ProbeResult* out, RowAcc* s_acc, CandQueue* s_queue, int& s_nchg,
   ChangeRec* outlog = nullptr, unsigned long long* outcursor = nullptr, long long outcap = 0)
{
   Ctrl* c = p.ctrl;
   const int tid = threadIdx.x;
   const int lane = tid & 31;

   if( c->poisoned )
   {
      if( tid == 0 )
      {
         out->status = GPULIN_PROBE_OVERFLOW;
         out->nrounds = 0;
         out->nchanges = 0;
         out->logoff = -1;
         out->nlog = 0;
      }
      return;
   }

   // ---- back to the node: the columns the previous probe changed are in its log
   if( restorevar >= 0 )
   {
      const unsigned long long n = min(c->logcount, (unsigned long long)c->logcap);
      for( unsigned long long i = tid; i <= n; i += PROBE_THREADS )
      {
         const int jj = i < n ? p.log[i].var : restorevar;
         const double2 b = base.bnd[jj];
         const_cast<double2*>(p.bnd)[jj] = b;
         reinterpret_cast<longlong2*>(p.cand)[jj] = reinterpret_cast<const longlong2*>(base.cand)[jj];
         noteBounds(p, jj, b.x, b.y);
      }
   }
   __syncthreads();

   // ---- start of the call (begin_kernel), the probed bound, the rows of its column
   if( tid == 0 )
   {
      c->maxrounds = maxrounds;
      c->logcap = logcap;
      c->round = 0;
      c->cont = 1;
      c->status = 0;
      c->cutoff = 0;
      c->ticket = 0;
      c->nchgcols = 0;
      c->nexact[0] = c->nexact[1] = c->nexact[2] = c->nexact[3] = 0;
      c->nsparse = 0;
      c->nfastrows = 0;
      c->logcount = 0;
      c->round_nchg = 0;
      for( int i = 0; i < NNZ_SLOTS; ++i )
         c->round_nnz[i] = 0;
      c->total_nchg = 0;
      c->total_nnz = 0;
      c->t_start = globaltimer();
      c->resume = 0;
      if( j >= 0 )
      {
         c->nmark[0][0] = c->nmark[0][1] = c->nmark[0][2] = 0;
         c->nmark[1][0] = c->nmark[1][1] = c->nmark[1][2] = 0;
         c->mb = 0;
         l += 0.0;
         u += 0.0;
         const_cast<double2*>(p.bnd)[j] = make_double2(l, u);
         reinterpret_cast<longlong2*>(p.cand)[j] = make_longlong2(~d2key(l), d2key(u));
         noteBounds(p, j, l, u);
         checkIntegralBounds(p, j, l, u);
      }
   }
   __syncthreads();
   bool overflow = false;
   if( j >= 0 )
      markColumnRows(p, j, tid, PROBE_THREADS);       // into mark list 1
   else
   {
      // rows left marked by an earlier call are not on the list that is read now: the general loop has to look
      const unsigned mb = c->mb;
      overflow = (c->nmark[mb][0] | c->nmark[mb][1] | c->nmark[mb][2]) != 0u;
   }
   __syncthreads();
   if( tid == 0 )
      c->mb ^= 1u;
   __syncthreads();

   for( ;; )
   {
      const unsigned mb = c->mb;
      const unsigned n0 = c->nmark[mb][0];
      const unsigned n1 = c->nmark[mb][1];
      const unsigned n2 = c->nmark[mb][2];
      if( c->cont == 0 || overflow )
         break;
      if( n0 > PROBE_MAXROWS || n1 > PROBE_MAXMEDIUM || n2 > 0u || c->logcount > (unsigned long long)c->logcap )
      {
         overflow = true;
         break;
      }
      if( tid == 0 )
         s_nchg = 0;
      const int* ml = p.marklist + (size_t)mb * 3 * MARKCAP;
      rangedListPhase(p, ml, n0, n1, n2, tid >> 5, PROBE_THREADS >> 5);
      exactPhase<true>(p, ml, n0, ml + MARKCAP, n1, ml + 2 * MARKCAP, n2, s_acc, s_queue, PROBE_THREADS);
      __syncthreads();
      int mychg = applyListPhase<SPARSE_G>(p, c->nchgcols, tid, PROBE_THREADS, c->round, c->logcap);
      mychg = __reduce_add_sync(0xffffffffu, mychg);
      if( lane == 0 && mychg != 0 )
         atomicAdd(&s_nchg, mychg);
      __syncthreads();
      if( tid == 0 )
         c->round_nchg = (unsigned long long)s_nchg;
      if( tid < 32 )
      {
         __syncwarp();
         controlStep<false>(c, 0);
      }
      __syncthreads();
   }

   if( keepmarks )
   {
      // a call on an ordinary handle: marked rows stay marked; after an overflow the general loop takes over
      if( tid == 0 )
      {
         if( overflow )
         {
            c->resume = 1;
            c->status = GPULIN_PROBE_OVERFLOW;
         }
         if( out != nullptr )
         {
            out->status = c->status;
            out->nrounds = c->round;
            out->nchanges = (long long)c->total_nchg;
         }
      }
      return;
   }
   // ---- whatever is still marked (cutoff, round limit, overflow) is unmarked: the next probe starts from the node
   {
      const unsigned mb = c->mb;
      const unsigned n0 = c->nmark[mb][0];
      const unsigned n1 = c->nmark[mb][1];
      const unsigned n2 = c->nmark[mb][2];
      const bool listed = n0 <= (unsigned)MARKCAP && n1 <= (unsigned)MARKCAP && n2 <= (unsigned)MARKCAP;
      if( listed )
      {
         const int* ml = p.marklist + (size_t)mb * 3 * MARKCAP;
         for( unsigned i = tid; i < n0; i += PROBE_THREADS )
            p.dirty[ml[i]] = ROW_CLEAN;
         for( unsigned i = tid; i < n1; i += PROBE_THREADS )
            p.dirty[ml[MARKCAP + i]] = ROW_CLEAN;
         for( unsigned i = tid; i < n2; i += PROBE_THREADS )
            p.dirty[ml[2 * MARKCAP + i]] = ROW_CLEAN;
      }
      __syncthreads();
      // ---- the bounds the probe implied (SCIPapplyProbingVar hands them out as proplbs / propubs, prop_probing.c:1203-1303):
      // ---- its change log goes to the output of the batch, one reservation per probe
      long long logoff = -1;
      const unsigned long long nl = min(c->logcount, (unsigned long long)c->logcap);
      if( outlog != nullptr )
      {
         __shared__ long long s_off;
         if( tid == 0 )
            s_off = (long long)atomicAdd(outcursor, nl);
         __syncthreads();
         if( s_off + (long long)nl <= outcap )
         {
            logoff = s_off;
            for( unsigned long long i = tid; i < nl; i += PROBE_THREADS )
               outlog[logoff + (long long)i] = p.log[i];
         }
      }
      if( tid == 0 )
      {
         c->nmark[mb][0] = c->nmark[mb][1] = c->nmark[mb][2] = 0;
         if( !listed || c->logcount > (unsigned long long)c->logcap )
            c->poisoned = 1;
         out->status = overflow ? GPULIN_PROBE_OVERFLOW : c->status;
         out->nrounds = c->round;
         out->nchanges = (long long)c->total_nchg;
         out->logoff = logoff;
         out->nlog = (long long)nl;
      }
   }
}

__global__ void __launch_bounds__(PROBE_THREADS, 1) probe_kernel(const DevProblem p, const DevProblem base, int restorevar, int j,
   double l, double u, int maxrounds, int logcap, int keepmarks, ProbeResult* out)
{
   __shared__ RowAcc s_acc[PROBE_THREADS / 32];
   __shared__ CandQueue s_queue[PROBE_THREADS / 32];
   __shared__ int s_nchg;
   probeBody(p, base, restorevar, j, l, u, maxrounds, logcap, keepmarks, out, s_acc, s_queue, s_nchg);
}

// the probes first, first + stride, ... < n of a batch, one after the other on this worker: one launch per worker and
// batch instead of one per probe (the host's launch rate was what bounded a batch)
__global__ void __launch_bounds__(PROBE_THREADS, 1) probe_list_kernel(const DevProblem p, const DevProblem base, int restorevar,
   const int* vars, const double* lbs, const double* ubs, int first, int stride, int n, int maxrounds, int logcap,
   ProbeResult* out, ChangeRec* outlog, unsigned long long* outcursor, long long outcap)
{
   __shared__ RowAcc s_acc[PROBE_THREADS / 32];
   __shared__ CandQueue s_queue[PROBE_THREADS / 32];
   __shared__ int s_nchg;
   for( int i = first; i < n; i += stride )
   {
      const int j = vars[i];
      probeBody(p, base, restorevar, j, lbs[i], ubs[i], maxrounds, logcap, 0, out + i, s_acc, s_queue, s_nchg, outlog, outcursor, outcap);
      restorevar = j;
      __syncthreads();
   }
}

// ---- bound (re)initialisation ----------------------------------------------------------------------------------
__global__ void set_bounds_kernel(const DevProblem p, const double* lb, const double* ub)
{
   const int stride = gridDim.x * blockDim.x;          // a multiple of 32: a warp owns the 32 columns of one word
   const int nallwords = max(p.nfreewords, (p.ncols + 31) / 32);
   for( int j0 = blockIdx.x * blockDim.x + (threadIdx.x & ~31); j0 < nallwords * 32; j0 += stride )
   {
      const int j = j0 + (threadIdx.x & 31);
      bool fr = false;
      if( j < p.ncols )
      {
         const double l = lb[j] + 0.0;
         const double u = ub[j] + 0.0;
         const_cast<double2*>(p.bnd)[j] = make_double2(l, u);
         reinterpret_cast<longlong2*>(p.cand)[j] = make_longlong2(~d2key(l), d2key(u));
         p.bndf[j] = midHalfWidth(l, u);
         p.colstate[j] = columnState(l, u);
         checkIntegralBounds(p, j, l, u);
         fr = isFree01(l, u);
      }
      const unsigned word = __ballot_sync(0xffffffffu, fr);
      if( (threadIdx.x & 31) == 0 )
         p.freebits[j0 >> 5] = word;
   }
   for( int w = blockIdx.x * blockDim.x + threadIdx.x; w < (p.ncols + 31) / 32; w += stride )
      p.colbits[w] = 0u;
   for( int r = blockIdx.x * blockDim.x + threadIdx.x; r < p.nrows; r += stride )
      p.dirty[r] = ROW_MARKED;
   for( int t = blockIdx.x * blockDim.x + threadIdx.x; t < p.ntiles; t += stride )
      p.tileflag[t] = 1;
}

// ---- bounds as 2 bits per column against resident reference bounds (gpulin_set_bounds_packed): at a node of a MIP almost
// ---- every column sits at its reference (global) bounds or is fixed to one of them -- 2 bits of information, not 16 bytes
constexpr unsigned BND_REF = 0u;        // the reference bounds
constexpr unsigned BND_FIXLB = 1u;      // fixed to the reference lower bound
constexpr unsigned BND_FIXUB = 2u;      // fixed to the reference upper bound
constexpr unsigned BND_EXPLICIT = 3u;   // on the explicit list (scattered by update_explicit_kernel afterwards)
__global__ void set_bounds_packed_kernel(const DevProblem p, const double2* ref, const unsigned* codes)
{
   const int stride = gridDim.x * blockDim.x;          // a multiple of 32: a warp owns the 32 columns of one word of freebits
   const int nallwords = max(p.nfreewords, (p.ncols + 31) / 32);
   for( int j0 = blockIdx.x * blockDim.x + (threadIdx.x & ~31); j0 < nallwords * 32; j0 += stride )
   {
      const int j = j0 + (threadIdx.x & 31);
      bool fr = false;
      if( j < p.ncols )
      {
         const unsigned code = (codes[j >> 4] >> (2 * (j & 15))) & 3u;
         const double2 r = ref[j];
         const double l = (code == BND_FIXUB ? r.y : r.x) + 0.0;
         const double u = (code == BND_FIXLB ? r.x : r.y) + 0.0;
         const_cast<double2*>(p.bnd)[j] = make_double2(l, u);
         reinterpret_cast<longlong2*>(p.cand)[j] = make_longlong2(~d2key(l), d2key(u));
         p.bndf[j] = midHalfWidth(l, u);
         p.colstate[j] = columnState(l, u);
         checkIntegralBounds(p, j, l, u);
         fr = isFree01(l, u);
      }
      const unsigned word = __ballot_sync(0xffffffffu, fr);
      if( (threadIdx.x & 31) == 0 )
         p.freebits[j0 >> 5] = word;
   }
   for( int w = blockIdx.x * blockDim.x + threadIdx.x; w < (p.ncols + 31) / 32; w += stride )
      p.colbits[w] = 0u;
   for( int r = blockIdx.x * blockDim.x + threadIdx.x; r < p.nrows; r += stride )
      p.dirty[r] = ROW_MARKED;
   for( int t = blockIdx.x * blockDim.x + threadIdx.x; t < p.ntiles; t += stride )
      p.tileflag[t] = 1;
}
__global__ void set_reference_kernel(int ncols, const double* lb, const double* ub, double2* ref)
{
   const int stride = gridDim.x * blockDim.x;
   for( int j = blockIdx.x * blockDim.x + threadIdx.x; j < ncols; j += stride )
      ref[j] = make_double2(lb[j] + 0.0, ub[j] + 0.0);
}
// the explicit entries of a packed bound vector (every row is marked already)
__global__ void update_explicit_kernel(const DevProblem p, long long nupd, const int* idx, const double* lb, const double* ub)
{
   const long long stride = (long long)gridDim.x * blockDim.x;
   for( long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nupd; i += stride )
   {
      const int j = idx[i];
      const double l = lb[i] + 0.0;
      const double u = ub[i] + 0.0;
      const_cast<double2*>(p.bnd)[j] = make_double2(l, u);
      reinterpret_cast<longlong2*>(p.cand)[j] = make_longlong2(~d2key(l), d2key(u));
      noteBounds(p, j, l, u);
      checkIntegralBounds(p, j, l, u);
   }
}
// the change log as 12-byte records: { column | is_upper << 31, new bound (2 x 32 bits) }; the round of an entry follows
// from its position (the log is round ordered, the per-round counts come with gpulin_get_round_stats)
__global__ void pack_log_kernel(const ChangeRec* log, long long n, unsigned* out)
{
   const long long stride = (long long)gridDim.x * blockDim.x;
   for( long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride )
   {
      const ChangeRec r = log[i];
      const unsigned long long b = (unsigned long long)__double_as_longlong(r.newbound);
      out[3 * i] = (unsigned)r.var | (r.is_upper ? 0x80000000u : 0u);
      out[3 * i + 1] = (unsigned)b;
      out[3 * i + 2] = (unsigned)(b >> 32);
   }
}

// the change log as ONE word per entry: column (29 bits) | code << 29 | is_upper << 31, code 0 / 1 = the new bound is 0 / 1
// (what a binary's bound can move to: 4 instead of 12 bytes over PCIe), code 2 = the bound is in the side list xout as
// { position of the entry in the log, low word, high word } -- the side list is in no particular order
constexpr unsigned CLOG_EXPLICIT = 2u;
__global__ void compact_log_kernel(const ChangeRec* log, long long n, unsigned* out, unsigned* xout, unsigned* xcount)
{
   const long long stride = (long long)gridDim.x * blockDim.x;
   for( long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride )
   {
      const ChangeRec r = log[i];
      const unsigned code = r.newbound == 0.0 ? 0u : (r.newbound == 1.0 ? 1u : CLOG_EXPLICIT);
      out[i] = (unsigned)r.var | (code << 29) | (r.is_upper ? 0x80000000u : 0u);
      if( code == CLOG_EXPLICIT )
      {
         const unsigned long long b = (unsigned long long)__double_as_longlong(r.newbound);
         const unsigned q = atomicAdd(xcount, 1u);
         xout[3 * (size_t)q] = (unsigned)i;
         xout[3 * (size_t)q + 1] = (unsigned)b;
         xout[3 * (size_t)q + 2] = (unsigned)(b >> 32);
      }
   }
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
      noteBounds(p, j, l, u);
      checkIntegralBounds(p, j, l, u);
      markColumnRows(p, j, 0, 1);
   }
}

// a handful of bound changes passed by value (no staging copy): a warp per column
struct SmallUpdate
{
   static constexpr int CAP = 8;
   int    n;
   int    idx[CAP];
   double lb[CAP];
   double ub[CAP];
};

__global__ void update_small_kernel(const DevProblem p, const SmallUpdate su)
{
   const int w = threadIdx.x >> 5;
   const int lane = threadIdx.x & 31;
   if( w >= su.n )
      return;
   const int j = su.idx[w];
   // two entries may name the same column: the later one wins, like in update_bounds_kernel's sequential semantics
   for( int k = w + 1; k < su.n; ++k )
   {
      if( su.idx[k] == j )
         return;
   }
   if( lane == 0 )
   {
      const double l = su.lb[w] + 0.0;
      const double u = su.ub[w] + 0.0;
      const_cast<double2*>(p.bnd)[j] = make_double2(l, u);
      reinterpret_cast<longlong2*>(p.cand)[j] = make_longlong2(~d2key(l), d2key(u));
      noteBounds(p, j, l, u);
      checkIntegralBounds(p, j, l, u);
   }
   markColumnRows(p, j, lane, 32);
}

// probing: one bound change passed by value (SCIPchgVarLbProbing / SCIPchgVarUbProbing), rows of the column marked
__global__ void update_one_kernel(const DevProblem p, int j, double l, double u)
{
   if( threadIdx.x == 0 )
   {
      l += 0.0;
      u += 0.0;
      const_cast<double2*>(p.bnd)[j] = make_double2(l, u);
      reinterpret_cast<longlong2*>(p.cand)[j] = make_longlong2(~d2key(l), d2key(u));
      noteBounds(p, j, l, u);
      checkIntegralBounds(p, j, l, u);
   }
   markColumnRows(p, j, threadIdx.x, blockDim.x);
}

// probing backtrack (SCIPbacktrackProbing): the columns the last probe changed -- they are in its change log -- and
// the probed column itself take the bounds of the node again; everything else was never touched
__global__ void restore_kernel(const DevProblem p, const DevProblem base, int probedvar)
{
   const unsigned long long n = min(p.ctrl->logcount, (unsigned long long)p.ctrl->logcap);
   const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
   for( unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += stride )
   {
      const int j = i < n ? p.log[i].var : probedvar;
      const double2 b = base.bnd[j];
      const_cast<double2*>(p.bnd)[j] = b;
      reinterpret_cast<longlong2*>(p.cand)[j] = reinterpret_cast<const longlong2*>(base.cand)[j];
      noteBounds(p, j, b.x, b.y);
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
   for( int t = blockIdx.x * blockDim.x + threadIdx.x; t < p.ntiles; t += stride )
      p.tileflag[t] = 1;
}

// start of a gpulin_propagate call: reset the loop state
__global__ void begin_kernel(Ctrl* c)
{
   if( c->resume )
   {
      // the call was started by probe_kernel (a small incremental call that outgrew its block): rounds, totals and the
      // change log go on; the marked rows are found through their flags
      c->resume = 0;
      c->cont = 1;
      c->status = 0;
      c->nmark[0][0] = c->nmark[0][1] = c->nmark[0][2] = 0;
      c->nmark[1][0] = c->nmark[1][1] = c->nmark[1][2] = 0;
      return;
   }
   c->round = 0;
   c->cont = 1;
   c->status = 0;
   c->cutoff = 0;
   c->ticket = 0;
   c->nchgcols = 0;
   c->nexact[0] = c->nexact[1] = c->nexact[2] = c->nexact[3] = 0;
   c->nmark[0][0] = c->nmark[0][1] = c->nmark[0][2] = 0;
   c->nmark[1][0] = c->nmark[1][1] = c->nmark[1][2] = 0;
   c->nsparse = 0;
   c->nfastrows = 0;
   c->logcount = 0;
   c->round_nchg = 0;
   for( int i = 0; i < NNZ_SLOTS; ++i )
      c->round_nnz[i] = 0;
   c->total_nchg = 0;
   c->total_nnz = 0;
   c->hist_push[0] = 0;
   c->ntrace = 0;
   c->t_start = globaltimer();
   traceStamp(c, TR_BEGIN);
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
