// gpulin_ranged.cuh -- ranged-row propagation (the gcd rule for equations and ranged rows) on the device.
//
// Restates rangedRowPropagation (cons_linear.c:5715-6696) with constraints/linear/rangedrowartcons = FALSE: the branches
// that ADD constraints (:6286-6306, :6565-6600, :6603-6688) are not bound propagation and are not built.  What is built:
// the partition of the unfixed variables of a row into a group of integer variables whose integral coefficients share a
// divisor gcd >= 2 and the rest (:5851-5957), the activity bounds of the rest (:5973-6018), the infeasibility test
// (:6036), the enumeration of the values the rest can take so that the row still has an integral solution (:6072-6130),
// and from it: a cutoff, a fixing, or two bounds of the one variable that can be bounded (:6136-6560).
//
// The reference walks the nonzeros of a row in storage order, and by the time the rule runs it has sorted the row
// (tightenBounds :7041 -> consdataSort -> consdataCompVarProp :3191).  gpulin_set_rangedrow keeps a copy of every ranged
// row in exactly that order (rrcols / rrvals): a warp reads it front to back, 32 nonzeros at a time, coalesced.
//
// One warp per row.  The only sequential part of the rule is the running gcd over the candidates of the first group (a
// candidate joins iff it shares a divisor with the gcd so far): it runs over the set bits of a ballot, uniformly in all
// lanes; everything else is a warp-parallel pass.  Sums are per-lane partial sums + a butterfly (the reference adds from
// the last nonzero to the first: for integer data -- the case in which the rule can conclude anything but a cutoff --
// both are exact).
#pragma once

#include "gpulin_device.cuh"

namespace gpl {

struct RangedRows
{
   int              n;          // ranged rows (two finite sides, at least three nonzeros); 0: the rule is off
   const long long* beg;        // n + 1: offsets into cols / vals
   const int*       row;        // n: the row (permuted numbering)
   const int*       cols;       // column | COL_INTEGRAL-style flag in bit 31, in the reference's sorted order
   const double*    vals;
   const int*       idx;        // nrows: index of a row in this structure, or -1
   unsigned*        scratch;    // per warp of a grid: membership bits of the second group, scratchwords words each
   int              scratchwords;
};

__device__ __forceinline__ double rrFloor(const Num& n, double x) { return floor(x + n.eps); }      // SCIPfloor
__device__ __forceinline__ double rrCeil(const Num& n, double x) { return ceil(x - n.eps); }        // SCIPceil
__device__ __forceinline__ bool rrIsInt(const Num& n, double x) { return x - rrFloor(n, x) <= n.eps; }   // SCIPisIntegral

__device__ __forceinline__ long long rrGcd(long long a, long long b)      // SCIPcalcGreComDiv (misc.c:9197): the value
{
   while( b != 0 )
   {
      const long long t = a % b;
      a = b;
      b = t;
   }
   return a;
}

struct RRElem      // one nonzero of the row that may have to be bounded
{
   double a, l, u;
   int    j;
   int    integral;
};

__device__ __forceinline__ RRElem rrShfl(const RRElem& e, int src)
{
   RRElem o;
   o.a = __shfl_sync(0xffffffffu, e.a, src);
   o.l = __shfl_sync(0xffffffffu, e.l, src);
   o.u = __shfl_sync(0xffffffffu, e.u, src);
   o.j = __shfl_sync(0xffffffffu, e.j, src);
   o.integral = __shfl_sync(0xffffffffu, e.integral, src);
   return o;
}

// the rule for ranged row number `rr`; all 32 lanes of a warp call it; `scratch`: this warp's membership words.
// Candidates go to the sink like those of the exact rules (forced: SCIPinferVarLbCons / UbCons / FixCons with force = TRUE)
__device__ __noinline__ void rangedRowWarp(const Num& n, const Sink& s, const RangedRows& R, const double2* bnd,
   const double2* sides, int rr, unsigned* scratch, int* cutoffflag)
{
   const int lane = threadIdx.x & 31;
   const long long beg = R.beg[rr];
   const int len = (int)(R.beg[rr + 1] - beg);
   const double2 sd = sides[R.row[rr]];
   const double feastol = n.feastol;

   // ---- pass 1, front to back: fixed activity, partition
   double fixedact = 0.0;
   int nfixed = 0;
   int ninf = 0;              // second group ("infcheckvars")
   int ncont = 0;
   int ncontbefore = 0;       // ... of them in front of the first candidate (the reference tests :5893 with that count)
   bool gcdisone = true;
   bool possiblegcd = true;
   bool started = false;      // the first unfixed candidate of the first group has been seen
   long long gcd = 0;
   RRElem firstinf;           // the first member of the second group ...
   RRElem starter;            // ... and the first of the first group: the only ones the rule can ever bound
   firstinf.a = starter.a = 0.0; firstinf.l = starter.l = 0.0; firstinf.u = starter.u = 0.0;
   firstinf.j = starter.j = -1; firstinf.integral = starter.integral = 0;
   for( int c0 = 0; c0 < len; c0 += 32 )
   {
      const int e = c0 + lane;
      RRElem x;
      x.a = 1.0; x.l = 0.0; x.u = 0.0; x.j = -1; x.integral = 0;
      bool valid = e < len;
      if( valid )
      {
         const int cw = R.cols[beg + e];
         x.a = R.vals[beg + e];
         x.j = cw & 0x7fffffff;
         x.integral = cw < 0 ? 1 : 0;
         const double2 b = bnd[x.j];
         x.l = b.x;
         x.u = b.y;
      }
      const bool fixed = valid && isEQ(n, x.l, x.u);
      const bool unfixed = valid && !fixed;
      // second-group type: not integral, or the coefficient is not integral, or it is +1 / -1 (:5865-5866)
      const bool second = !x.integral || !rrIsInt(n, x.a) || isEQ(n, fabs(x.a), 1.0);
      if( fixed )
         fixedact += x.l * x.a;
      nfixed += __popc(__ballot_sync(0xffffffffu, fixed));
      bool member = unfixed && second;                  // of the second group
      const unsigned candm = __ballot_sync(0xffffffffu, unfixed && !second);
      if( !started )
      {
         const unsigned before = candm != 0u ? ((1u << (__ffs(candm) - 1)) - 1u) : 0xffffffffu;
         ncontbefore += __popc(__ballot_sync(0xffffffffu, unfixed && !x.integral) & before);
      }
      // the running gcd over the candidates, in order (uniform in all lanes)
      unsigned m = candm;
      const long long myabs = (long long)(fabs(x.a) + feastol);
      while( m != 0u )
      {
         const int k = __ffs(m) - 1;
         m &= m - 1u;
         const long long ak = __shfl_sync(0xffffffffu, myabs, k);
         if( !started )
         {
            started = true;
            gcd = ak;
            const RRElem st = rrShfl(x, k);
            starter = st;
         }
         else
         {
            const long long g = rrGcd(gcd, ak);
            if( g == 1 )
            {
               if( lane == k )
                  member = true;                        // shares no divisor: second group (:5940-5949)
            }
            else
               gcd = g;
         }
      }
      // a second-type variable takes the flags along; a candidate that fell out of the first group does not (:5925-5929 vs :5942)
      const unsigned typem = __ballot_sync(0xffffffffu, unfixed && second);
      const unsigned contm = __ballot_sync(0xffffffffu, unfixed && !x.integral);
      const unsigned notonem = __ballot_sync(0xffffffffu, unfixed && second && !isEQ(n, fabs(x.a), 1.0));
      if( typem != 0u )
         possiblegcd = false;
      if( notonem != 0u )
         gcdisone = false;
      ncont += __popc(contm);
      const unsigned memberm = __ballot_sync(0xffffffffu, member);
      if( ninf == 0 && memberm != 0u )
         firstinf = rrShfl(x, __ffs(memberm) - 1);
      ninf += __popc(memberm);
      if( lane == 0 )
         scratch[c0 >> 5] = memberm;
   }
   __syncwarp();
#pragma unroll
   for( int d = 16; d >= 1; d >>= 1 )
      fixedact += __shfl_xor_sync(0xffffffffu, fixedact, d);

   if( isHuge(n, fabs(fixedact)) )      // :5819
      return;
   const double lhs = sd.x - fixedact;
   const double rhs = sd.y - fixedact;
   const int nunfixed = len - nfixed;
   if( !started || ncontbefore + 2 > nunfixed || ninf == 0 )      // :5889, :5893, :5962
      return;

   // ---- pass 2: activity bounds and the gcd of the second group (:5973-6018, :6054-6066)
   double minact = 0.0;
   double maxact = 0.0;
   bool invalid = false;
   long long gcdinf = 0;
   for( int c0 = 0; c0 < len; c0 += 32 )
   {
      const unsigned memberm = scratch[c0 >> 5];
      if( ((memberm >> lane) & 1u) != 0u )
      {
         const int e = c0 + lane;
         const double a = R.vals[beg + e];
         const double2 b = bnd[R.cols[beg + e] & 0x7fffffff];
         if( isInf(n, -b.x) || isInf(n, b.y) )
            invalid = true;       // (whichever activity it makes infinite: the rule gives up, :6016)
         else
         {
            const double al = a * b.x;
            const double au = a * b.y;
            if( a < 0.0 )
            {
               maxact += al;
               minact += au;
            }
            else
            {
               minact += al;
               maxact += au;
            }
         }
         gcdinf = rrGcd(gcdinf, (long long)(fabs(a) + feastol));
      }
   }
#pragma unroll
   for( int d = 16; d >= 1; d >>= 1 )
   {
      minact += __shfl_xor_sync(0xffffffffu, minact, d);
      maxact += __shfl_xor_sync(0xffffffffu, maxact, d);
      gcdinf = rrGcd(gcdinf, __shfl_xor_sync(0xffffffffu, gcdinf, d));
   }
   if( __any_sync(0xffffffffu, invalid) || isHuge(n, -minact) || isHuge(n, maxact) )
      return;

   const double dg = (double)gcd;
   // no multiple of the gcd between the sides (:6036-6047)
   if( !rrIsInt(n, (lhs - maxact) / dg) && isGT(n, rrCeil(n, (lhs - maxact) / dg) * dg, rhs - minact) )
   {
      if( lane == 0 )
         *cutoffflag = 1;
      return;
   }
   if( ncont != 0 )
      return;
   long long gcdinfvars = -1;
   if( possiblegcd )
      gcdinfvars = gcdinf;
   else if( gcdisone )
      gcdinfvars = 1;
   if( gcdinfvars < 1 )
      return;

   // ---- the values the second group can take so that the row keeps an integral solution (:6072-6099), 32 at a time.  The
   // ---- pattern repeats after gcd / gcd(gcd, gcdinfvars) steps: three periods hold three solutions if there is one
   const double dgi = (double)gcdinfvars;
   const long long period = gcd / rrGcd(gcd, gcdinfvars);
   const long long maxsteps = 3 * period + 3;
   double minvalue = 0.0;
   double maxvalue = 0.0;
   int nsols = 0;
   {
      const double v0 = rrCeil(n, minact - feastol);
      bool done = false;
      for( long long t0 = 0; t0 < maxsteps && !done; t0 += 32 )
      {
         const double value = v0 + (double)(t0 + lane) * dgi;
         const bool inrange = isLE(n, value, maxact);
         double value2 = value + dg * rrCeil(n, (lhs - value) / dg);
         if( !isGE(n, value2, lhs) )
            value2 += dg;
         const bool sol = inrange && isLE(n, value2, rhs);
         unsigned sm = __ballot_sync(0xffffffffu, sol);
         const unsigned rm = __ballot_sync(0xffffffffu, inrange);
         while( sm != 0u && nsols < 3 )
         {
            const int k = __ffs(sm) - 1;
            sm &= sm - 1u;
            ++nsols;
            if( nsols == 3 )
               break;
            const double vk = v0 + (double)(t0 + k) * dgi;
            if( nsols == 1 )
               minvalue = vk;
            maxvalue = vk;
         }
         if( nsols == 3 || rm != 0xffffffffu )
            done = true;
      }
   }
   // more than two: the last one from above (:6103-6130)
   if( nsols == 3 )
   {
      const double v1 = rrFloor(n, maxact + feastol);
      bool done = false;
      for( long long t0 = 0; t0 < maxsteps && !done; t0 += 32 )
      {
         const double value = v1 - (double)(t0 + lane) * dgi;
         const bool inrange = isGE(n, value, minact);
         double value2 = value + dg * rrFloor(n, (rhs - value) / dg);
         if( !isLE(n, value2, rhs) )
            value2 -= dg;
         const bool sol = inrange && isGE(n, value2, lhs);
         const unsigned sm = __ballot_sync(0xffffffffu, sol);
         const unsigned rm = __ballot_sync(0xffffffffu, inrange);
         if( sm != 0u )
         {
            maxvalue = v1 - (double)(t0 + __ffs(sm) - 1) * dgi;
            done = true;
         }
         else if( rm != 0xffffffffu )
            done = true;
      }
   }
   if( nsols == 0 )      // :6136
   {
      if( lane == 0 )
         *cutoffflag = 1;
      return;
   }

   // ---- the one variable that can be bounded: the only member of the second group (:6156, :6323), or the only unfixed
   // ---- variable outside it (:6188, :6385) -- then the first group is its first candidate alone
   RRElem tg;
   bool insecond;
   if( ninf == 1 )
   {
      tg = firstinf;
      insecond = true;
   }
   else if( ninf == nunfixed - 1 )
   {
      tg = starter;
      insecond = false;
   }
   else
      return;
   if( lane != 0 )
      return;
   bool cutoff = false;
   bool touched = false;
   if( nsols == 1 )
   {
      // SCIPinferVarFixCons (scip_var.c:6896): the lower bound, then the upper bound, both forced
      double fix;
      if( insecond )
         fix = maxvalue / tg.a;                                                                       // :6172
      else
         fix = tg.a < 0.0 ? rrFloor(n, (lhs - maxvalue) / tg.a) : rrCeil(n, (lhs - maxvalue) / tg.a);   // :6224-6231
      inferLb(n, s, tg.j, tg.integral != 0, fix, tg.l, tg.u, true, cutoff, touched);
      if( !cutoff )
         inferUb(n, s, tg.j, tg.integral != 0, fix, tg.l, tg.u, true, cutoff, touched);
   }
   else
   {
      double nlb;
      double nub;
      if( insecond )
      {
         nlb = tg.a < 0.0 ? maxvalue / tg.a : minvalue / tg.a;        // :6332-6341
         nub = tg.a < 0.0 ? minvalue / tg.a : maxvalue / tg.a;
      }
      else if( tg.a < 0.0 )
      {
         nlb = rrFloor(n, (rhs - minvalue) / tg.a);                   // :6424-6427
         nub = rrFloor(n, (lhs - maxvalue) / tg.a);
      }
      else
      {
         nlb = rrCeil(n, (lhs - maxvalue) / tg.a);                    // :6431-6432
         nub = rrCeil(n, (rhs - minvalue) / tg.a);
      }
      if( nlb > tg.l )
         inferLb(n, s, tg.j, tg.integral != 0, nlb, tg.l, tg.u, true, cutoff, touched);
      if( !cutoff && nub < tg.u )
         inferUb(n, s, tg.j, tg.integral != 0, nub, tg.l, tg.u, true, cutoff, touched);
   }
   if( cutoff )
      *cutoffflag = 1;
   if( touched )
   {
      const bool first = raiseColumnBit(s, tg.j);
      listChangedColumn(s, tg.j, first);
   }
}

} // namespace gpl
