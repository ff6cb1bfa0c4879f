// gpulin_device.cuh -- device-side arithmetic of the linear bound propagation path (sm_100a).
//
// Everything here is fp64 + small integer counters; there is no GEMM-shaped work and no tensor-core use.
// The functions restate the reference's rules (file:line cited per function, relative to
// /root/reference/src/scip) as straight-line device code.  Compiled with -fmad=false so that every product
// and sum rounds exactly like the reference's C code (built with -ffp-contract=off).
#pragma once

#include <cstdint>
#include <cuda_runtime.h>
#include <cooperative_groups.h>

namespace gpl {

struct Num
{
   double inf;        // SCIPinfinity          def.h:172
   double eps;        // SCIPepsilon           def.h:173
   double sumeps;     // SCIPsumepsilon        def.h:174
   double feastol;    // SCIPfeastol           def.h:175
   double bstreps;    // numerics/boundstreps  def.h:180
   double huge;       // SCIPgetHugeValue      def.h:184
   double maxeasy;    // constraints/linear/maxeasyactivitydelta  cons_linear.c:135
};

// ---- order-preserving int64 encoding of fp64 bounds: atomicMax/atomicMin on the keys is
// ---- max/min on the doubles; "+ 0.0" folds -0.0 into +0.0; an involution, so decode == encode
__device__ __forceinline__ long long d2key(double x)
{
   long long u = __double_as_longlong(x + 0.0);
   return u >= 0 ? u : (u ^ 0x7fffffffffffffffLL);
}
__device__ __forceinline__ double key2d(long long k)
{
   return __longlong_as_double(k >= 0 ? k : (k ^ 0x7fffffffffffffffLL));
}

// ---- tolerance predicates: set.c:6824-6976, 7133-7150, 7254-7372; misc.c:11162 -----------------------------
__device__ __forceinline__ bool isInf(const Num& n, double v) { return v >= n.inf; }
__device__ __forceinline__ bool isHuge(const Num& n, double v) { return v >= n.huge; }
__device__ __forceinline__ bool isLT(const Num& n, double a, double b) { return a - b < -n.eps; }
__device__ __forceinline__ bool isLE(const Num& n, double a, double b) { return a - b <= n.eps; }
__device__ __forceinline__ bool isGT(const Num& n, double a, double b) { return a - b > n.eps; }
__device__ __forceinline__ bool isGE(const Num& n, double a, double b) { return a - b >= -n.eps; }
__device__ __forceinline__ bool isEQ(const Num& n, double a, double b) { return fabs(a - b) <= n.eps; }
// plain maximum / minimum of two doubles that are never NaN here (fmax/fmin cost ~8 instructions each for their NaN rules)
__device__ __forceinline__ double dmax(double a, double b) { return a > b ? a : b; }
__device__ __forceinline__ double dmin(double a, double b) { return a < b ? a : b; }
__device__ __forceinline__ double relDiff(double a, double b)
{
   double q = dmax(1.0, dmax(fabs(a), fabs(b)));
   return (a - b) / q;
}
// relDiff(a,b) < -feastol  /  > feastol  (set.c:7254-7372).  The fp64 division is only executed when the product form
// cannot decide: d/q and feastol are compared with a guard band of 1e-10 relative, far above the rounding of either side.
__device__ __forceinline__ bool isFeasLT(const Num& n, double a, double b)
{
   const double q = dmax(1.0, dmax(fabs(a), fabs(b)));
   const double d = a - b;
   const double t = n.feastol * q;
   if( d < -t * 1.0000000001 )
      return true;
   if( d > -t * 0.9999999999 )
      return false;
   return d / q < -n.feastol;
}
__device__ __forceinline__ bool isFeasGT(const Num& n, double a, double b)
{
   const double q = dmax(1.0, dmax(fabs(a), fabs(b)));
   const double d = a - b;
   const double t = n.feastol * q;
   if( d > t * 1.0000000001 )
      return true;
   if( d < t * 0.9999999999 )
      return false;
   return d / q > n.feastol;
}

// set.c:7711-7753
__device__ __forceinline__ bool isLbBetter(const Num& n, double newlb, double oldlb, double oldub)
{
   if( oldlb < 0.0 && newlb >= 0.0 )
      return true;
   double m = dmax(dmin(oldub - oldlb, fabs(oldlb)), 1e-3);
   return newlb - oldlb > n.bstreps * m;
}
__device__ __forceinline__ bool isUbBetter(const Num& n, double newub, double oldlb, double oldub)
{
   if( oldub > 0.0 && newub <= 0.0 )
      return true;
   double m = dmax(dmin(oldub - oldlb, fabs(oldub)), 1e-3);
   return newub - oldub < -(n.bstreps * m);
}

// var.c:1957-1974 / 2006-2023
__device__ __forceinline__ double adjustedLb(const Num& n, bool integral, double lb)
{
   if( lb < 0.0 && isInf(n, -lb) )
      return -n.inf;
   if( lb > 0.0 && isInf(n, lb) )
      return n.inf;
   if( integral )
      return ceil(lb - n.feastol);
   if( lb > 0.0 && lb < n.eps )
      return 0.0;
   return lb;
}
__device__ __forceinline__ double adjustedUb(const Num& n, bool integral, double ub)
{
   if( ub > 0.0 && isInf(n, ub) )
      return n.inf;
   if( ub < 0.0 && isInf(n, -ub) )
      return -n.inf;
   if( integral )
      return floor(ub + n.feastol);
   if( ub < 0.0 && ub > -n.eps )
      return 0.0;
   return ub;
}

// ---- double-double accumulation: dbldblarith.h:154-187 (SCIPdbldblSum21) ------------------------------------
__device__ __forceinline__ void dd_add(double& hi, double& lo, double b)
{
   double s = hi + b;
   double t = s - hi;
   double e = (hi - (s - t)) + (b - t);
   lo = e + lo;
   hi = s;
}
// (hi,lo) += (bhi,blo): merges the partial sums of two lanes
__device__ __forceinline__ void dd_add_dd(double& hi, double& lo, double bhi, double blo)
{
   double s = hi + bhi;
   double t = s - hi;
   double e = (hi - (s - t)) + (bhi - t);
   lo = (e + lo) + blo;
   hi = s;
}

// ---- per-row activity state --------------------------------------------------------------------------------
// The eight counters of infinite / huge contributions (cons_linear.c:187-266) are kept as 8-bit fields of one 64-bit
// word that saturate at 127.  The infinity counters are only ever asked "0, 1 or more" (canTightenBounds :5234,
// getMinActivity :2373-2392, residual counters :2726-2805); the huge counters enter the relaxed activities of the row
// verdict as count * hugeval (:2386, :2481), so they keep their true value (a row with more than 127 contributions of
// 1e15 or more each has an activity beyond 1e17 whichever count is used).
enum CntField { MINPOSINF = 0, MINNEGINF = 1, MINPOSHUGE = 2, MINNEGHUGE = 3,
                MAXPOSINF = 4, MAXNEGINF = 5, MAXPOSHUGE = 6, MAXNEGHUGE = 7 };
typedef unsigned long long Cnt;
constexpr unsigned CNT_SAT = 127u;

__device__ __forceinline__ unsigned cntGet(Cnt cnt, int f) { return (unsigned)(cnt >> (8 * f)) & 0xFFu; }
__device__ __forceinline__ void cntInc(Cnt& cnt, int f)
{
   if( cntGet(cnt, f) < CNT_SAT )
      cnt += 1ull << (8 * f);
}
__device__ __forceinline__ Cnt cntDec(Cnt cnt, int f) { return cnt - (1ull << (8 * f)); }
__device__ __forceinline__ Cnt cntMerge(Cnt a, Cnt b)
{
   const Cnt x = a + b;                                      // fields 0..254: no carry between the bytes
   const Cnt ge = x & 0x8080808080808080ull;                 // bit 7 where a field is >= 128
   return (x | (ge - (ge >> 7))) & 0x7f7f7f7f7f7f7f7full;    // ... those become 127
}
__device__ __forceinline__ unsigned cntMinSum(Cnt c) { return cntGet(c, 0) + cntGet(c, 1) + cntGet(c, 2) + cntGet(c, 3); }
__device__ __forceinline__ unsigned cntMaxSum(Cnt c) { return cntGet(c, 4) + cntGet(c, 5) + cntGet(c, 6) + cntGet(c, 7); }

struct RowAcc
{
   double   minhi, minlo;   // finite part of the minimal activity (double-double)
   double   maxhi, maxlo;   // finite part of the maximal activity
   double   maxdelta;       // max_i |a_i| (ub_i - lb_i), infinity if a variable is unbounded (:1542-1599)
   Cnt      cnt;
};

__device__ __forceinline__ void accInit(RowAcc& r)
{
   r.minhi = r.minlo = r.maxhi = r.maxlo = 0.0;
   r.maxdelta = 0.0;
   r.cnt = 0ull;
}

// one bound of one term enters one activity: consdataUpdateActivities with oldbound = 0 (:1773-1948)
__device__ __forceinline__ void addContributionSlow(const Num& n, double a, double bound, double& hi, double& lo,
   Cnt& cnt, int fposinf, int fneginf, int fposhuge, int fneghuge)
{
   if( isInf(n, fabs(bound)) )
      cntInc(cnt, bound > 0.0 ? fposinf : fneginf);
   else
   {
      double c = a * bound;
      if( isHuge(n, fabs(c)) )
         cntInc(cnt, c > 0.0 ? fposhuge : fneghuge);
      else
         dd_add(hi, lo, c);
   }
}

// the rare part of accElem: an infinite bound or a huge product
__device__ __noinline__ void accElemSlow(const Num& n, RowAcc& r, double a, double l, double u)
{
   if( a > 0.0 )
   {
      addContributionSlow(n, a, l, r.minhi, r.minlo, r.cnt, MINPOSINF, MINNEGINF, MINPOSHUGE, MINNEGHUGE);
      addContributionSlow(n, a, u, r.maxhi, r.maxlo, r.cnt, MAXPOSINF, MAXNEGINF, MAXPOSHUGE, MAXNEGHUGE);
   }
   else
   {
      // negative coefficient: the infinity counters are switched, the huge counters follow the sign of the
      // contribution (:1737-1769)
      addContributionSlow(n, a, l, r.maxhi, r.maxlo, r.cnt, MAXNEGINF, MAXPOSINF, MAXPOSHUGE, MAXNEGHUGE);
      addContributionSlow(n, a, u, r.minhi, r.minlo, r.cnt, MINNEGINF, MINPOSINF, MINPOSHUGE, MINNEGHUGE);
   }
   if( isInf(n, -l) || isInf(n, u) )
      r.maxdelta = n.inf;
   else
      r.maxdelta = dmax(r.maxdelta, fabs(a) * (u - l));
}

// classification + accumulation of one nonzero (a, [l,u]); the common case (finite bounds, no huge product)
// is branch-free
__device__ __forceinline__ void accElem(const Num& n, RowAcc& r, double a, double l, double u)
{
   const bool pos = a > 0.0;
   const double bmin = pos ? l : u;
   const double bmax = pos ? u : l;
   const double cmin = a * bmin;
   const double cmax = a * bmax;
   const bool fast = (fabs(l) < n.inf) && (fabs(u) < n.inf) && (fabs(cmin) < n.huge) && (fabs(cmax) < n.huge);
   if( fast )
   {
      dd_add(r.minhi, r.minlo, cmin);
      dd_add(r.maxhi, r.maxlo, cmax);
      r.maxdelta = dmax(r.maxdelta, fabs(a) * (u - l));
   }
   else
   {
      // through a copy: the accumulator itself never has its address taken and stays in registers (passing `r` would
      // pin it to local memory: a load and a store around every nonzero)
      RowAcc t = r;
      accElemSlow(n, t, a, l, u);
      r = t;
   }
}

__device__ __forceinline__ void accMerge(RowAcc& r, const RowAcc& o)
{
   dd_add_dd(r.minhi, r.minlo, o.minhi, o.minlo);
   dd_add_dd(r.maxhi, r.maxlo, o.maxhi, o.maxlo);
   r.maxdelta = dmax(r.maxdelta, o.maxdelta);
   r.cnt = cntMerge(r.cnt, o.cnt);
}

// getMinActivity / getMaxActivity: cons_linear.c:2345-2434 / 2440-2529
__device__ __forceinline__ void getMinActivity(const Num& n, double hi, double lo, unsigned posinf, unsigned neginf,
   unsigned poshuge, unsigned neghuge, double delta, bool goodrelax, double& act, bool& tight, bool& settoinf)
{
   if( neginf > 0 )
   {
      act = -n.inf; settoinf = true; tight = (posinf == 0);
   }
   else if( posinf > 0 )
   {
      act = n.inf; settoinf = true; tight = true;
   }
   else if( neghuge > 0 || (poshuge > 0 && !goodrelax) )
   {
      act = -n.inf; settoinf = true; tight = false;
   }
   else
   {
      dd_add(hi, lo, -delta);
      if( poshuge > 0 )
      {
         dd_add(hi, lo, (double)poshuge * n.huge);
         tight = false;
      }
      else
         tight = true;
      act = hi + lo;
      settoinf = false;
   }
}
__device__ __forceinline__ void getMaxActivity(const Num& n, double hi, double lo, unsigned posinf, unsigned neginf,
   unsigned poshuge, unsigned neghuge, double delta, bool goodrelax, double& act, bool& tight, bool& settoinf)
{
   if( posinf > 0 )
   {
      act = n.inf; settoinf = true; tight = (neginf == 0);
   }
   else if( neginf > 0 )
   {
      act = -n.inf; settoinf = true; tight = true;
   }
   else if( poshuge > 0 || (neghuge > 0 && !goodrelax) )
   {
      act = n.inf; settoinf = true; tight = false;
   }
   else
   {
      dd_add(hi, lo, -delta);
      if( neghuge > 0 )
      {
         dd_add(hi, lo, -(double)neghuge * n.huge);
         tight = false;
      }
      else
         tight = true;
      act = hi + lo;
      settoinf = false;
   }
}

// ---- what the candidate phase needs to know about its row ------------------------------------------------------
struct RowInfo
{
   RowAcc acc;
   double lhs, rhs;
   double minact, maxact;   // finite parts as doubles (easy path: cons_linear.c:5450-5515)
   double slackR, slackL;   // clamped rhs - minact / maxact - lhs of the easy path
   bool   easy;             // maxactdelta < maxeasyactivitydelta  (:7086)
   bool   force;            // single-variable row                 (:7029)
   bool   rhsfin, lhsfin;
};

// row gates of tightenBounds (cons_linear.c:7021-7086); returns true if the row's nonzeros must be visited.
// *cutoff is set when the easy path's side test proves infeasibility (:5450, :5506).
__device__ __forceinline__ bool rowGates(const Num& n, RowInfo& ri, int len, bool& cutoff)
{
   const RowAcc& a = ri.acc;
   ri.force = (len == 1);
   ri.rhsfin = !isInf(n, ri.rhs);
   ri.lhsfin = !isInf(n, -ri.lhs);
   // canTightenBounds (:5214-5238)
   if( cntMinSum(a.cnt) > 1u && cntMaxSum(a.cnt) > 1u )
      return false;
   // all variables fixed (:7057)
   if( fabs(a.maxdelta) <= n.feastol )
      return false;
   if( !isInf(n, a.maxdelta) )
   {
      double minact, maxact;
      bool t1, t2, s1, s2;
      getMinActivity(n, a.minhi, a.minlo, cntGet(a.cnt, MINPOSINF), cntGet(a.cnt, MINNEGINF), cntGet(a.cnt, MINPOSHUGE),
         cntGet(a.cnt, MINNEGHUGE), 0.0, false, minact, t1, s1);
      getMaxActivity(n, a.maxhi, a.maxlo, cntGet(a.cnt, MAXPOSINF), cntGet(a.cnt, MAXNEGINF), cntGet(a.cnt, MAXPOSHUGE),
         cntGet(a.cnt, MAXNEGHUGE), 0.0, false, maxact, t2, s2);
      double slack = (!ri.rhsfin || s1) ? n.inf : (ri.rhs - minact);
      double surplus = (!ri.lhsfin || s2) ? n.inf : (maxact - ri.lhs);
      if( isLE(n, a.maxdelta, dmin(slack, surplus)) )
         return false;
   }
   ri.easy = isLT(n, a.maxdelta, n.maxeasy);
   if( ri.easy )
   {
      ri.minact = a.minhi + a.minlo;
      ri.maxact = a.maxhi + a.maxlo;
      ri.slackR = 0.0;
      ri.slackL = 0.0;
      if( ri.rhsfin )
      {
         if( isFeasLT(n, ri.rhs, ri.minact) )
            cutoff = true;
         double s = ri.rhs - ri.minact;
         ri.slackR = (s > n.eps) ? s : 0.0;
      }
      if( ri.lhsfin )
      {
         if( isFeasLT(n, ri.maxact, ri.lhs) )
            cutoff = true;
         double s = ri.maxact - ri.lhs;
         ri.slackL = (s > n.eps) ? s : 0.0;
      }
   }
   return true;
}

// row verdict of propagateCons (cons_linear.c:7715-7742), goodrelax = TRUE
__device__ __forceinline__ bool rowInfeasible(const Num& n, const RowAcc& a, double lhs, double rhs)
{
   double minact, maxact;
   bool t1, t2, s1, s2;
   getMinActivity(n, a.minhi, a.minlo, cntGet(a.cnt, MINPOSINF), cntGet(a.cnt, MINNEGINF), cntGet(a.cnt, MINPOSHUGE),
      cntGet(a.cnt, MINNEGHUGE), 0.0, true, minact, t1, s1);
   getMaxActivity(n, a.maxhi, a.maxlo, cntGet(a.cnt, MAXPOSINF), cntGet(a.cnt, MAXNEGINF), cntGet(a.cnt, MAXPOSHUGE),
      cntGet(a.cnt, MAXNEGHUGE), 0.0, true, maxact, t2, s2);
   return isFeasGT(n, minact, rhs) || isFeasLT(n, maxact, lhs);
}

// redundancy verdict of propagateCons (cons_linear.c:7743-7753): a row that is not infeasible is redundant for the current
// bounds iff GE(minactivity, lhs) and LE(maxactivity, rhs), goodrelax = TRUE -- the reference then deletes it locally
__device__ __forceinline__ bool rowRedundant(const Num& n, const RowAcc& a, double lhs, double rhs)
{
   double minact, maxact;
   bool t1, t2, s1, s2;
   getMinActivity(n, a.minhi, a.minlo, cntGet(a.cnt, MINPOSINF), cntGet(a.cnt, MINNEGINF), cntGet(a.cnt, MINPOSHUGE),
      cntGet(a.cnt, MINNEGHUGE), 0.0, true, minact, t1, s1);
   getMaxActivity(n, a.maxhi, a.maxlo, cntGet(a.cnt, MAXPOSINF), cntGet(a.cnt, MAXNEGINF), cntGet(a.cnt, MAXPOSHUGE),
      cntGet(a.cnt, MAXNEGHUGE), 0.0, true, maxact, t2, s2);
   if( isFeasGT(n, minact, rhs) || isFeasLT(n, maxact, lhs) )
      return false;
   return isGE(n, minact, lhs) && isLE(n, maxact, rhs);
}

// ---- commit filter: SCIPinferVarUbCons/LbCons (scip_var.c:7071-7157 / 6965-7051) followed by the last drop of
// ---- SCIPnodeAddBoundinfer (tree.c:2020-2059), judged against the round-start bounds [l,u]; a surviving value
// ---- is merged with atomicMin on its int64 key.  Candidate layout per column j (one 16-byte pair):
// ----    cand[2j]   = ~d2key(lb)   (bitwise NOT reverses the order: a larger lower bound is a smaller key)
// ----    cand[2j+1] =  d2key(ub)
// ---- so both sides tighten by MIN -- the ranks' candidates merge with one MIN per key (atomicMin of the packed
// ---- exchange, or one ncclMin all-reduce over the whole vector).
// Several GPUs of one node (see gpulin_kernels.cuh, "dense rounds sharded over the GPUs of a node"): every rank commits
// into its OWN key vector; after the exact kernel of a dense round the touched columns travel once, packed, into the
// inbox of every other rank (peer memory over NVLink), and every rank merges what it received with atomicMin.
constexpr int MAX_PEERS = 8;
// An inbox holds, per parity of the exchange number and per source rank, one 8-byte header word and `cap` entries of 32
// bytes.  Everything is written with plain stores through peer memory and validates itself -- no fence, no flag that has
// to be ordered behind the data (the scheme of NCCL's low-latency protocols):
//    header = exchange number << 32 | verdict << 31 | number of entries        (one 8-byte store: atomic)
//    entry  = { column, exchange number, ~key(lb) } { key(ub), exchange number, 0 }      (two 16-byte stores; the reader
//             spins until both halves carry the number it waits for)
struct PeerEntry
{
   uint4 a;     // x = column, y = exchange number, (z, w) = ~key(lb)
   uint4 b;     // (x, y) = key(ub), z = exchange number, w = 0
};
struct PeerHeader
{
   unsigned long long word;
   unsigned long long nnz;       // nonzeros the source swept in this round (statistics only: not validated)
};
constexpr size_t PEER_HDR_BYTES = 2 * MAX_PEERS * sizeof(PeerHeader);
struct PeerTable
{
   int            n;                  // ranks
   int            rank;               // this rank
   long long      cap;                // entries per (parity, source): the number of columns
   unsigned char* box[MAX_PEERS];     // inbox of every rank ([rank] = the local one): headers | entries
};
__host__ __device__ __forceinline__ size_t peerBoxBytes(int nranks, long long cap)
{
   return PEER_HDR_BYTES + 2 * (size_t)nranks * (size_t)cap * sizeof(PeerEntry);
}
__device__ __forceinline__ PeerHeader* peerHeader(unsigned char* box, int parity, int src)
{
   return reinterpret_cast<PeerHeader*>(box) + parity * MAX_PEERS + src;
}
__device__ __forceinline__ PeerEntry* peerEntries(const PeerTable& t, unsigned char* box, int parity, int src)
{
   return reinterpret_cast<PeerEntry*>(box + PEER_HDR_BYTES) + ((size_t)parity * t.n + src) * (size_t)t.cap;
}

struct Sink
{
   long long*       cand;      // 2*ncols candidate keys
   unsigned*        colbits;   // one bit per column: "a key moved this round"
   int*             chglist;   // the columns whose bit was raised this round, in no particular order
   unsigned*        nchgcols;  // length of chglist
   bool             listed;    // small rounds: the first candidate that reaches a column appends it to chglist (bit test with
                               // an answer, list counter: two dependent round trips).  Dense rounds: the bit goes up with a
                               // reduction that nobody waits for, and collect_kernel turns the bits into the list afterwards
};

// a reduction: nobody waits for its answer (a candidate that gets here beats the round-start bound, so its column changes
// this round whichever candidate wins)
__device__ __forceinline__ bool commitKey(const Sink& s, size_t idx, long long key)
{
   atomicMin(&s.cand[idx], key);
   return true;
}

// the first candidate of a round that reaches a column puts it on the list the apply kernel works through.  Split in
// two so that a thread can have the bit tests of several columns in flight before it looks at the first answer.
__device__ __forceinline__ bool raiseColumnBit(const Sink& s, int j)
{
   const unsigned bit = 1u << (j & 31);
   if( !s.listed )
   {
      atomicOr(&s.colbits[j >> 5], bit);
      return false;
   }
   return (atomicOr(&s.colbits[j >> 5], bit) & bit) == 0u;
}
__device__ __forceinline__ void listChangedColumn(const Sink& s, int j, bool first)
{
   if( first )
   {
      // one atomic on the list counter per group of lanes that got here together
      const cooperative_groups::coalesced_group g = cooperative_groups::coalesced_threads();
      unsigned pos = 0u;
      if( g.thread_rank() == 0 )
         pos = atomicAdd(s.nchgcols, (unsigned)g.size());
      pos = g.shfl(pos, 0) + g.thread_rank();
      s.chglist[pos] = j;
   }
}

__device__ __forceinline__ void inferUb(const Num& n, const Sink& s, int j, bool integral, double newub, double l,
   double u, bool force, bool& cutoff, bool& touched)
{
   newub = adjustedUb(n, integral, newub);
   if( isInf(n, -newub) || isFeasLT(n, newub, l) )
   {
      cutoff = true;
      return;
   }
   newub = dmax(newub, l);
   if( force ? isGE(n, newub, u) : !isUbBetter(n, newub, l, u) )
      return;
   if( !isLT(n, newub, u) )
      return;
   // every value that gets here is below the round-start bound, so the column changes this round whichever
   // candidate wins: no need to wait for the atomic's result
   if( commitKey(s, 2 * (size_t)j + 1, d2key(newub)) )
      touched = true;
}
__device__ __forceinline__ void inferLb(const Num& n, const Sink& s, int j, bool integral, double newlb, double l,
   double u, bool force, bool& cutoff, bool& touched)
{
   newlb = adjustedLb(n, integral, newlb);
   if( isInf(n, newlb) || isFeasGT(n, newlb, u) )
   {
      cutoff = true;
      return;
   }
   newlb = dmin(newlb, u);
   if( force ? isLE(n, newlb, l) : !isLbBetter(n, newlb, l, u) )
      return;
   if( !isGT(n, newlb, l) )
      return;
   if( commitKey(s, 2 * (size_t)j, ~d2key(newlb)) )
      touched = true;
}

// tightenVarUb / tightenVarLb: cons_linear.c:5242-5307 / 5311-5376
__device__ __forceinline__ void tightenVarUb(const Num& n, const Sink& s, int j, bool integral, double newub, double l,
   double u, bool force, bool& cutoff, bool& touched)
{
   newub = adjustedUb(n, integral, newub);
   if( force || isUbBetter(n, newub, l, u) )
      inferUb(n, s, j, integral, newub, l, u, force, cutoff, touched);
}
__device__ __forceinline__ void tightenVarLb(const Num& n, const Sink& s, int j, bool integral, double newlb, double l,
   double u, bool force, bool& cutoff, bool& touched)
{
   newlb = adjustedLb(n, integral, newlb);
   if( force || isLbBetter(n, newlb, l, u) )
      inferLb(n, s, j, integral, newlb, l, u, force, cutoff, touched);
}

// candidate bounds of one nonzero: tightenVarBoundsEasy (cons_linear.c:5380-5653) or tightenVarBounds (:6700-6974)
__device__ __forceinline__ void candidates(const Num& n, const Sink& s, const RowInfo& ri, double a, int j, bool integral,
   double l, double u, bool& cutoff, bool& touched)
{
   const bool pos = a > 0.0;
   if( ri.easy )
   {
      const double alpha = pos ? a * (u - l) : a * (l - u);
      if( ri.rhsfin && ((alpha - ri.slackR > n.sumeps) || (ri.force && alpha - ri.slackR > n.eps)) )
      {
         if( pos )
            tightenVarUb(n, s, j, integral, l + (ri.slackR / a), l, u, ri.force, cutoff, touched);
         else
            tightenVarLb(n, s, j, integral, u + ri.slackR / a, l, u, ri.force, cutoff, touched);
      }
      if( ri.lhsfin && ((alpha - ri.slackL > n.sumeps) || (ri.force && alpha - ri.slackL > n.eps)) )
      {
         if( pos )
            tightenVarLb(n, s, j, integral, u - (ri.slackL / a), l, u, ri.force, cutoff, touched);
         else
            tightenVarUb(n, s, j, integral, l - (ri.slackL / a), l, u, ri.force, cutoff, touched);
      }
      return;
   }

   // general case via residual activities: consdataGetActivityResiduals (cons_linear.c:2661-2806), goodrelax = FALSE
   const RowAcc& r = ri.acc;
   const double minactbound = pos ? l : -u;
   const double maxactbound = pos ? u : -l;
   const double absval = fabs(a);
   double minres, maxres;
   bool mintight, maxtight, minsettoinf, maxsettoinf;
   {
      Cnt c = r.cnt;
      double delta = 0.0;
      if( isInf(n, minactbound) ) c = cntDec(c, MINPOSINF);
      else if( isInf(n, -minactbound) ) c = cntDec(c, MINNEGINF);
      else if( isHuge(n, minactbound * absval) ) c = cntDec(c, MINPOSHUGE);
      else if( isHuge(n, -minactbound * absval) ) c = cntDec(c, MINNEGHUGE);
      else delta = absval * minactbound;
      getMinActivity(n, r.minhi, r.minlo, cntGet(c, MINPOSINF), cntGet(c, MINNEGINF), cntGet(c, MINPOSHUGE),
         cntGet(c, MINNEGHUGE), delta, false, minres, mintight, minsettoinf);
   }
   {
      Cnt c = r.cnt;
      double delta = 0.0;
      if( isInf(n, -maxactbound) ) c = cntDec(c, MAXNEGINF);
      else if( isInf(n, maxactbound) ) c = cntDec(c, MAXPOSINF);
      else if( isHuge(n, absval * maxactbound) ) c = cntDec(c, MAXPOSHUGE);
      else if( isHuge(n, -absval * maxactbound) ) c = cntDec(c, MAXNEGHUGE);
      else delta = absval * maxactbound;
      getMaxActivity(n, r.maxhi, r.maxlo, cntGet(c, MAXPOSINF), cntGet(c, MAXNEGINF), cntGet(c, MAXPOSHUGE),
         cntGet(c, MAXNEGHUGE), delta, false, maxres, maxtight, maxsettoinf);
   }
   if( !minsettoinf && ri.rhsfin && mintight && (isLT(n, fabs(ri.rhs), 1.0) || !isEQ(n, minres / ri.rhs, 1.0)) )
   {
      const double nb = (ri.rhs - minres) / a;
      if( pos )
      {
         if( !isInf(n, nb) && ((ri.force && isLT(n, nb, u)) || (integral && isFeasLT(n, nb, u)) || isUbBetter(n, nb, l, u)) )
            inferUb(n, s, j, integral, nb, l, u, ri.force, cutoff, touched);
      }
      else
      {
         if( !isInf(n, -nb) && ((ri.force && isGT(n, nb, l)) || (integral && isFeasGT(n, nb, l)) || isLbBetter(n, nb, l, u)) )
            inferLb(n, s, j, integral, nb, l, u, ri.force, cutoff, touched);
      }
   }
   if( !maxsettoinf && ri.lhsfin && maxtight && (isLT(n, fabs(ri.lhs), 1.0) || !isEQ(n, maxres / ri.lhs, 1.0)) )
   {
      const double nb = (ri.lhs - maxres) / a;
      if( pos )
      {
         if( !isInf(n, -nb) && ((ri.force && isGT(n, nb, l)) || (integral && isFeasGT(n, nb, l)) || isLbBetter(n, nb, l, u)) )
            inferLb(n, s, j, integral, nb, l, u, ri.force, cutoff, touched);
      }
      else
      {
         if( !isInf(n, nb) && ((ri.force && isLT(n, nb, u)) || (integral && isFeasLT(n, nb, u)) || isUbBetter(n, nb, l, u)) )
            inferUb(n, s, j, integral, nb, l, u, ri.force, cutoff, touched);
      }
   }
}

} // namespace gpl
