/* linprop_oracle.c -- TEST INFRASTRUCTURE ONLY (see linprop_oracle.h).
 *
 * Restates, function by function, the reference path (all citations are /root/reference/src/scip/...):
 *   term classification + activities   cons_linear.c:1604-1992 (consdataUpdateActivities), :2283-2339
 *   maxactdelta                        cons_linear.c:1542-1599
 *   getMinActivity / getMaxActivity    cons_linear.c:2345-2529
 *   residual activities                cons_linear.c:2661-2806
 *   canTightenBounds                   cons_linear.c:5214-5238
 *   tightenVarUb / tightenVarLb        cons_linear.c:5242-5376
 *   tightenVarBoundsEasy               cons_linear.c:5380-5653
 *   tightenVarBounds                   cons_linear.c:6700-6974
 *   tightenBounds (row gates)          cons_linear.c:6980-7154
 *   propagateCons (row verdict)        cons_linear.c:7715-7754
 *   adjustedLb / adjustedUb            var.c:1957-2023
 *   SCIPinferVarLbCons / UbCons        scip_var.c:6965-7157
 *   SCIPnodeAddBoundinfer (last drop)  tree.c:2020-2059
 *   tolerance predicates               set.c:6824-6976, 7133-7150, 7254-7372, 7433-7455, 7711-7753; misc.c:11162
 *   double-double sum                  dbldblarith.h:154-187
 *   rangedRowPropagation (gcd rule)    cons_linear.c:5715-6696  (optional: oracle_propagate_ranged)
 *   consdataCompVarProp (row order)    cons_linear.c:3191-3257  (the gcd rule walks the row in this order)
 *
 * Deliberate differences from the reference's control flow:
 *   - rounds are synchronous (Jacobi): every candidate of a round is judged against the round-start bounds and
 *     the best accepted candidate per variable bound wins; the reference is sequential and event driven;
 *   - activities are recomputed from scratch in double-double each round (the reference updates them
 *     incrementally); the "unreliable update" recomputation (cons_linear.c:2580-2657) is therefore vacuous;
 *   - the binary-skip of sorted rows (:7134), MAXTIGHTENROUNDS (:6976) and the boundstightened flags only
 *     save work in the reference and are not restated;
 *   - a variable whose new bounds cross by less than feastol within one round gets lb := ub (the reference's
 *     result depends on processing order there).
 */
#include "linprop_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct
{
   double hi;
   double lo;
} DD;

/* dbldblarith.h:154-187: SCIPdbldblSum21 */
static void dd_sum21(DD* r, DD a, double b)
{
   double s = a.hi + b;
   double t = s - a.hi;
   double e = (a.hi - (s - t)) + (b - t);
   r->lo = e + a.lo;
   r->hi = s;
}

void oracle_dd_sum21(double* rhi, double* rlo, double ahi, double alo, double b)
{
   DD a;
   DD r;
   a.hi = ahi;
   a.lo = alo;
   dd_sum21(&r, a, b);
   *rhi = r.hi;
   *rlo = r.lo;
}

#define DD_TO_DBL(x) ((x).hi + (x).lo)

void oracle_default_numerics(ORACLE_NUMERICS* num)
{
   num->infinity = 1e20;
   num->epsilon = 1e-9;
   num->sumepsilon = 1e-6;
   num->feastol = 1e-6;
   num->boundstreps = 0.05;
   num->hugeval = 1e15;
   num->maxeasyactivitydelta = 1e6;
}

/* --- tolerance predicates (set.c) --- */
static int isInf(const ORACLE_NUMERICS* n, double v) { return v >= n->infinity; }
static int isHuge(const ORACLE_NUMERICS* n, double v) { return v >= n->hugeval; }
static int isLT(const ORACLE_NUMERICS* n, double a, double b) { return a - b < -n->epsilon; }
static int isLE(const ORACLE_NUMERICS* n, double a, double b) { return a - b <= n->epsilon; }
static int isGT(const ORACLE_NUMERICS* n, double a, double b) { return a - b > n->epsilon; }
static int isGE(const ORACLE_NUMERICS* n, double a, double b) { return a - b >= -n->epsilon; }
static int isEQ(const ORACLE_NUMERICS* n, double a, double b) { return fabs(a - b) <= n->epsilon; }
static int isPositive(const ORACLE_NUMERICS* n, double a) { return a > n->epsilon; }
static int isSumGT(const ORACLE_NUMERICS* n, double a, double b) { return a - b > n->sumepsilon; }

/* misc.c:11162 */
static double relDiff(double a, double b)
{
   double q = 1.0;
   double absa = fabs(a);
   double absb = fabs(b);
   if( absa > q ) q = absa;
   if( absb > q ) q = absb;
   return (a - b) / q;
}
static int isFeasLT(const ORACLE_NUMERICS* n, double a, double b) { return relDiff(a, b) < -n->feastol; }
static int isFeasGT(const ORACLE_NUMERICS* n, double a, double b) { return relDiff(a, b) > n->feastol; }
static int isFeasZero(const ORACLE_NUMERICS* n, double a) { return fabs(a) <= n->feastol; }

/* set.c:7711-7753 */
static int isLbBetter(const ORACLE_NUMERICS* n, double newlb, double oldlb, double oldub)
{
   double m;
   if( oldlb < 0.0 && newlb >= 0.0 )
      return 1;
   m = fabs(oldlb);
   if( oldub - oldlb < m ) m = oldub - oldlb;
   if( m < 1e-3 ) m = 1e-3;
   return newlb - oldlb > n->boundstreps * m;
}
static int isUbBetter(const ORACLE_NUMERICS* n, double newub, double oldlb, double oldub)
{
   double m;
   if( oldub > 0.0 && newub <= 0.0 )
      return 1;
   m = fabs(oldub);
   if( oldub - oldlb < m ) m = oldub - oldlb;
   if( m < 1e-3 ) m = 1e-3;
   return newub - oldub < -(n->boundstreps * m);
}

/* var.c:1957-1974 / 2006-2023 */
static double adjustedLb(const ORACLE_NUMERICS* n, int integral, double lb)
{
   if( lb < 0.0 && isInf(n, -lb) )
      return -n->infinity;
   else if( lb > 0.0 && isInf(n, lb) )
      return n->infinity;
   else if( integral )
      return ceil(lb - n->feastol);
   else if( lb > 0.0 && lb < n->epsilon )
      return 0.0;
   return lb;
}
static double adjustedUb(const ORACLE_NUMERICS* n, int integral, double ub)
{
   if( ub > 0.0 && isInf(n, ub) )
      return n->infinity;
   else if( ub < 0.0 && isInf(n, -ub) )
      return -n->infinity;
   else if( integral )
      return floor(ub + n->feastol);
   else if( ub < 0.0 && ub > -n->epsilon )
      return 0.0;
   return ub;
}

/* --- row activities --- */
typedef struct
{
   DD     minact;
   DD     maxact;
   int    minposinf, minneginf, minposhuge, minneghuge;
   int    maxposinf, maxneginf, maxposhuge, maxneghuge;
   double maxactdelta;
} ROWACT;

/* one bound of one term enters one activity: cons_linear.c:1773-1948 with oldbound = 0 */
static void addContribution(const ORACLE_NUMERICS* n, double a, double bound, DD* act, int* posinf, int* neginf,
   int* poshuge, int* neghuge)
{
   if( isInf(n, fabs(bound)) )
   {
      if( bound > 0.0 )
         (*posinf)++;
      else
         (*neginf)++;
   }
   else
   {
      double c = a * bound;
      if( isHuge(n, fabs(c)) )
      {
         if( c > 0.0 )
            (*poshuge)++;
         else
            (*neghuge)++;
      }
      else if( c != 0.0 )
         dd_sum21(act, *act, c);
   }
}

static void rowActivities(const ORACLE_PROBLEM* p, const ORACLE_NUMERICS* n, const double* lb, const double* ub,
   int64_t r, ROWACT* ra)
{
   int64_t k;
   memset(ra, 0, sizeof(*ra));
   for( k = p->rowptr[r]; k < p->rowptr[r + 1]; ++k )
   {
      double a = p->vals[k];
      double l = lb[p->colidx[k]];
      double u = ub[p->colidx[k]];
      if( a > 0.0 )
      {
         /* lower bound + pos. coef -> minactivity; upper bound + pos. coef -> maxactivity (:1725-1759) */
         addContribution(n, a, l, &ra->minact, &ra->minposinf, &ra->minneginf, &ra->minposhuge, &ra->minneghuge);
         addContribution(n, a, u, &ra->maxact, &ra->maxposinf, &ra->maxneginf, &ra->maxposhuge, &ra->maxneghuge);
      }
      else
      {
         /* lower bound + neg. coef -> maxactivity, upper bound + neg. coef -> minactivity; the infinity counters
          * are switched, the huge counters follow the sign of the contribution (:1737-1769) */
         addContribution(n, a, l, &ra->maxact, &ra->maxneginf, &ra->maxposinf, &ra->maxposhuge, &ra->maxneghuge);
         addContribution(n, a, u, &ra->minact, &ra->minneginf, &ra->minposinf, &ra->minposhuge, &ra->minneghuge);
      }
   }
   /* cons_linear.c:1574-1598 */
   ra->maxactdelta = 0.0;
   for( k = p->rowptr[r + 1] - 1; k >= p->rowptr[r]; --k )
   {
      double l = lb[p->colidx[k]];
      double u = ub[p->colidx[k]];
      double delta;
      if( isInf(n, -l) || isInf(n, u) )
      {
         ra->maxactdelta = n->infinity;
         break;
      }
      delta = fabs(p->vals[k]) * (u - l);
      if( delta > ra->maxactdelta )
         ra->maxactdelta = delta;
   }
}

/* cons_linear.c:2345-2434 */
static void getMinActivity(const ORACLE_NUMERICS* n, DD finite, int posinf, int neginf, int poshuge, int neghuge,
   double delta, int goodrelax, double* minactivity, int* istight, int* issettoinfinity)
{
   if( neginf > 0 )
   {
      *minactivity = -n->infinity;
      *issettoinfinity = 1;
      *istight = (posinf == 0);
   }
   else if( posinf > 0 )
   {
      *minactivity = n->infinity;
      *issettoinfinity = 1;
      *istight = 1;
   }
   else if( neghuge > 0 || (poshuge > 0 && !goodrelax) )
   {
      *minactivity = -n->infinity;
      *issettoinfinity = 1;
      *istight = 0;
   }
   else
   {
      DD tmp;
      dd_sum21(&tmp, finite, -delta);
      if( poshuge > 0 )
      {
         dd_sum21(&tmp, tmp, poshuge * n->hugeval);
         *istight = 0;
      }
      else
         *istight = 1;
      *minactivity = DD_TO_DBL(tmp);
      *issettoinfinity = 0;
   }
}

/* cons_linear.c:2440-2529 */
static void getMaxActivity(const ORACLE_NUMERICS* n, DD finite, int posinf, int neginf, int poshuge, int neghuge,
   double delta, int goodrelax, double* maxactivity, int* istight, int* issettoinfinity)
{
   if( posinf > 0 )
   {
      *maxactivity = n->infinity;
      *issettoinfinity = 1;
      *istight = (neginf == 0);
   }
   else if( neginf > 0 )
   {
      *maxactivity = -n->infinity;
      *issettoinfinity = 1;
      *istight = 1;
   }
   else if( poshuge > 0 || (neghuge > 0 && !goodrelax) )
   {
      *maxactivity = n->infinity;
      *issettoinfinity = 1;
      *istight = 0;
   }
   else
   {
      DD tmp;
      dd_sum21(&tmp, finite, -delta);
      if( neghuge > 0 )
      {
         dd_sum21(&tmp, tmp, -neghuge * n->hugeval);
         *istight = 0;
      }
      else
         *istight = 1;
      *maxactivity = DD_TO_DBL(tmp);
      *issettoinfinity = 0;
   }
}

/* cons_linear.c:2661-2806 (goodrelax = FALSE as at :6749) */
static void getActivityResiduals(const ORACLE_NUMERICS* n, const ROWACT* ra, double a, double l, double u,
   double* minres, double* maxres, int* mintight, int* maxtight, int* minsettoinf, int* maxsettoinf)
{
   double minactbound;
   double maxactbound;
   double absval;

   if( a > 0.0 )
   {
      minactbound = l;
      maxactbound = u;
      absval = a;
   }
   else
   {
      minactbound = -u;
      maxactbound = -l;
      absval = -a;
   }

   if( isInf(n, minactbound) )
      getMinActivity(n, ra->minact, ra->minposinf - 1, ra->minneginf, ra->minposhuge, ra->minneghuge, 0.0, 0, minres, mintight, minsettoinf);
   else if( isInf(n, -minactbound) )
      getMinActivity(n, ra->minact, ra->minposinf, ra->minneginf - 1, ra->minposhuge, ra->minneghuge, 0.0, 0, minres, mintight, minsettoinf);
   else if( isHuge(n, minactbound * absval) )
      getMinActivity(n, ra->minact, ra->minposinf, ra->minneginf, ra->minposhuge - 1, ra->minneghuge, 0.0, 0, minres, mintight, minsettoinf);
   else if( isHuge(n, -minactbound * absval) )
      getMinActivity(n, ra->minact, ra->minposinf, ra->minneginf, ra->minposhuge, ra->minneghuge - 1, 0.0, 0, minres, mintight, minsettoinf);
   else
      getMinActivity(n, ra->minact, ra->minposinf, ra->minneginf, ra->minposhuge, ra->minneghuge, absval * minactbound, 0, minres, mintight, minsettoinf);

   if( isInf(n, -maxactbound) )
      getMaxActivity(n, ra->maxact, ra->maxposinf, ra->maxneginf - 1, ra->maxposhuge, ra->maxneghuge, 0.0, 0, maxres, maxtight, maxsettoinf);
   else if( isInf(n, maxactbound) )
      getMaxActivity(n, ra->maxact, ra->maxposinf - 1, ra->maxneginf, ra->maxposhuge, ra->maxneghuge, 0.0, 0, maxres, maxtight, maxsettoinf);
   else if( isHuge(n, absval * maxactbound) )
      getMaxActivity(n, ra->maxact, ra->maxposinf, ra->maxneginf, ra->maxposhuge - 1, ra->maxneghuge, 0.0, 0, maxres, maxtight, maxsettoinf);
   else if( isHuge(n, -absval * maxactbound) )
      getMaxActivity(n, ra->maxact, ra->maxposinf, ra->maxneginf, ra->maxposhuge, ra->maxneghuge - 1, 0.0, 0, maxres, maxtight, maxsettoinf);
   else
      getMaxActivity(n, ra->maxact, ra->maxposinf, ra->maxneginf, ra->maxposhuge, ra->maxneghuge, absval * maxactbound, 0, maxres, maxtight, maxsettoinf);
}

/* --- commit filter: SCIPinferVarUbCons (scip_var.c:7071-7157) + SCIPnodeAddBoundinfer (tree.c:2020-2059),
 *     judged against the round-start bounds [l,u]; the surviving value is min-merged into *newub --- */
static void inferUb(const ORACLE_NUMERICS* n, int integral, double newub, double l, double u, int force,
   double* outub, int* cutoff)
{
   newub = adjustedUb(n, integral, newub);
   if( isInf(n, -newub) || isFeasLT(n, newub, l) )
   {
      *cutoff = 1;
      return;
   }
   if( newub < l )
      newub = l;
   if( (force && isGE(n, newub, u)) || (!force && !isUbBetter(n, newub, l, u)) )
      return;
   /* tree.c: adjust again (idempotent), clamp, ignore if not LT */
   newub = adjustedUb(n, integral, newub);
   if( newub < l )
      newub = l;
   if( !isLT(n, newub, u) )
      return;
   if( newub < *outub )
      *outub = newub;
}

static void inferLb(const ORACLE_NUMERICS* n, int integral, double newlb, double l, double u, int force,
   double* outlb, int* cutoff)
{
   newlb = adjustedLb(n, integral, newlb);
   if( isInf(n, newlb) || isFeasGT(n, newlb, u) )
   {
      *cutoff = 1;
      return;
   }
   if( newlb > u )
      newlb = u;
   if( (force && isLE(n, newlb, l)) || (!force && !isLbBetter(n, newlb, l, u)) )
      return;
   newlb = adjustedLb(n, integral, newlb);
   if( newlb > u )
      newlb = u;
   if( !isGT(n, newlb, l) )
      return;
   if( newlb > *outlb )
      *outlb = newlb;
}

/* cons_linear.c:5242-5307 / 5311-5376 */
static void tightenVarUb(const ORACLE_NUMERICS* n, int integral, double newub, double l, double u, int force,
   double* outub, int* cutoff)
{
   newub = adjustedUb(n, integral, newub);
   if( force || isUbBetter(n, newub, l, u) )
      inferUb(n, integral, newub, l, u, force, outub, cutoff);
}
static void tightenVarLb(const ORACLE_NUMERICS* n, int integral, double newlb, double l, double u, int force,
   double* outlb, int* cutoff)
{
   newlb = adjustedLb(n, integral, newlb);
   if( force || isLbBetter(n, newlb, l, u) )
      inferLb(n, integral, newlb, l, u, force, outlb, cutoff);
}

/* cons_linear.c:5380-5653; returns 1 on cutoff */
static int tightenVarBoundsEasy(const ORACLE_NUMERICS* n, const ROWACT* ra, double a, double lhs, double rhs,
   int integral, double l, double u, int force, double* outlb, double* outub)
{
   int cutoff = 0;
   double minact = DD_TO_DBL(ra->minact);
   double maxact = DD_TO_DBL(ra->maxact);

   if( !isInf(n, rhs) )
   {
      double slack;
      double alpha;
      if( isFeasLT(n, rhs, minact) )
         return 1;
      slack = rhs - minact;
      if( !isPositive(n, slack) )
         slack = 0.0;
      alpha = (a > 0.0) ? a * (u - l) : a * (l - u);
      if( isSumGT(n, alpha, slack) || (force && isGT(n, alpha, slack)) )
      {
         if( a > 0.0 )
            tightenVarUb(n, integral, l + (slack / a), l, u, force, outub, &cutoff);
         else
            tightenVarLb(n, integral, u + slack / a, l, u, force, outlb, &cutoff);
         if( cutoff )
            return 1;
      }
   }
   if( !isInf(n, -lhs) )
   {
      double slack;
      double alpha;
      if( isFeasLT(n, maxact, lhs) )
         return 1;
      slack = maxact - lhs;
      if( !isPositive(n, slack) )
         slack = 0.0;
      alpha = (a > 0.0) ? a * (u - l) : a * (l - u);
      if( isSumGT(n, alpha, slack) || (force && isGT(n, alpha, slack)) )
      {
         if( a > 0.0 )
            tightenVarLb(n, integral, u - (slack / a), l, u, force, outlb, &cutoff);
         else
            tightenVarUb(n, integral, l - (slack / a), l, u, force, outub, &cutoff);
         if( cutoff )
            return 1;
      }
   }
   return 0;
}

/* cons_linear.c:6700-6974; returns 1 on cutoff */
static int tightenVarBounds(const ORACLE_NUMERICS* n, const ROWACT* ra, double a, double lhs, double rhs,
   int integral, double l, double u, int force, double* outlb, double* outub)
{
   double minres, maxres;
   int mintight, maxtight, minsettoinf, maxsettoinf;
   int cutoff = 0;

   getActivityResiduals(n, ra, a, l, u, &minres, &maxres, &mintight, &maxtight, &minsettoinf, &maxsettoinf);

   if( !minsettoinf && !isInf(n, rhs) && mintight && (isLT(n, fabs(rhs), 1.0) || !isEQ(n, minres / rhs, 1.0)) )
   {
      double nb = (rhs - minres) / a;
      if( a > 0.0 )
      {
         if( !isInf(n, nb) && ((force && isLT(n, nb, u)) || (integral && isFeasLT(n, nb, u)) || isUbBetter(n, nb, l, u)) )
            inferUb(n, integral, nb, l, u, force, outub, &cutoff);
      }
      else
      {
         if( !isInf(n, -nb) && ((force && isGT(n, nb, l)) || (integral && isFeasGT(n, nb, l)) || isLbBetter(n, nb, l, u)) )
            inferLb(n, integral, nb, l, u, force, outlb, &cutoff);
      }
      if( cutoff )
         return 1;
   }
   if( !maxsettoinf && !isInf(n, -lhs) && maxtight && (isLT(n, fabs(lhs), 1.0) || !isEQ(n, maxres / lhs, 1.0)) )
   {
      double nb = (lhs - maxres) / a;
      if( a > 0.0 )
      {
         if( !isInf(n, -nb) && ((force && isGT(n, nb, l)) || (integral && isFeasGT(n, nb, l)) || isLbBetter(n, nb, l, u)) )
            inferLb(n, integral, nb, l, u, force, outlb, &cutoff);
      }
      else
      {
         if( !isInf(n, nb) && ((force && isLT(n, nb, u)) || (integral && isFeasLT(n, nb, u)) || isUbBetter(n, nb, l, u)) )
            inferUb(n, integral, nb, l, u, force, outub, &cutoff);
      }
      if( cutoff )
         return 1;
   }
   return 0;
}

/* ---------------------------------------------------------------------------------------------------------------
 * ranged-row propagation: rangedRowPropagation, cons_linear.c:5715-6696, with constraints/linear/rangedrowartcons = FALSE
 * (the branches that ADD constraints, :6286-6306, :6565-6600, :6603-6688, are not part of the bound propagation path).
 *
 * The rule walks the nonzeros of the row in storage order, and by the time it runs the reference has sorted the row
 * (tightenBounds :7041 -> consdataSort :3337 -> consdataCompVarProp :3191): binaries first by decreasing |a|, then the
 * other integers by decreasing |a (ub_global - lb_global)|, then the continuous variables; ties by SCIPvarGetProbindex.
 * RANGED holds that order for every row with two finite sides and at least three nonzeros.
 * --------------------------------------------------------------------------------------------------------------- */
typedef struct
{
   const int64_t* ord;       /* nnz: ord[rowptr[r] + v] = position (into colidx / vals) of the v-th nonzero of row r in sorted order */
} RANGED;

static double epsFloor(const ORACLE_NUMERICS* n, double x) { return floor(x + n->epsilon); }       /* SCIPfloor, def.h:197 */
static double epsCeil(const ORACLE_NUMERICS* n, double x) { return ceil(x - n->epsilon); }         /* SCIPceil,  def.h:198 */
static int epsIsInt(const ORACLE_NUMERICS* n, double x) { return x - epsFloor(n, x) <= n->epsilon; } /* SCIPisIntegral, def.h:203 */

static long long calcGcd(long long a, long long b)      /* SCIPcalcGreComDiv (misc.c:9197): the value, by Euclid */
{
   while( b != 0 )
   {
      long long t = a % b;
      a = b;
      b = t;
   }
   return a;
}

/* comparator context (qsort has no user pointer in C99) */
static const ORACLE_PROBLEM* g_sortprob = NULL;
static const double* g_sortglb = NULL;
static const double* g_sortgub = NULL;
static const int32_t* g_sorttie = NULL;

static int isBinaryCol(int64_t j)
{
   /* SCIPvarIsBinary (var.c:23801): binary type, or integral with global bounds inside [0,1] */
   return g_sortprob->vartype[j] != 0 && g_sortglb[j] >= 0.0 && g_sortgub[j] <= 1.0;
}

static int compVarProp(const void* x, const void* y)     /* consdataCompVarProp, cons_linear.c:3191-3257 */
{
   const int64_t k1 = *(const int64_t*)x;
   const int64_t k2 = *(const int64_t*)y;
   const int64_t j1 = g_sortprob->colidx[k1];
   const int64_t j2 = g_sortprob->colidx[k2];
   const int b1 = isBinaryCol(j1);
   const int b2 = isBinaryCol(j2);
   const long long t1 = g_sorttie != NULL ? g_sorttie[j1] : j1;
   const long long t2 = g_sorttie != NULL ? g_sorttie[j2] : j2;
   if( b1 != b2 )
      return b1 ? -1 : +1;
   if( b1 )
   {
      const double a1 = fabs(g_sortprob->vals[k1]);
      const double a2 = fabs(g_sortprob->vals[k2]);
      if( a1 - a2 > 1e-9 )
         return -1;
      if( a2 - a1 > 1e-9 )
         return +1;
      return t1 < t2 ? -1 : (t1 > t2 ? +1 : 0);
   }
   else
   {
      const int i1 = g_sortprob->vartype[j1] != 0;
      const int i2 = g_sortprob->vartype[j2] != 0;
      if( i1 != i2 )
         return i1 ? -1 : +1;          /* integer type before continuous */
      if( !i1 )
         return t1 < t2 ? -1 : (t1 > t2 ? +1 : 0);
      else
      {
         const double c1 = fabs(g_sortprob->vals[k1] * (g_sortgub[j1] - g_sortglb[j1]));
         const double c2 = fabs(g_sortprob->vals[k2] * (g_sortgub[j2] - g_sortglb[j2]));
         if( c1 - c2 > 1e-9 )
            return -1;
         if( c2 - c1 > 1e-9 )
            return +1;
         return t1 < t2 ? -1 : (t1 > t2 ? +1 : 0);
      }
   }
}

static int isRangedRow(const ORACLE_PROBLEM* p, const ORACLE_NUMERICS* n, int64_t r)
{
   return p->rowptr[r + 1] - p->rowptr[r] >= 3 && !isInf(n, -p->lhs[r]) && !isInf(n, p->rhs[r]);
}

/* one row; candidates are merged into newlb / newub (judged against the round-start bounds lb / ub); returns 1 on cutoff */
static int rangedRowPropagation(const ORACLE_PROBLEM* p, const ORACLE_NUMERICS* n, const RANGED* rr, const double* lb,
   const double* ub, int64_t r, double* newlb, double* newub)
{
   const int64_t beg = p->rowptr[r];
   const int nvars = (int)(p->rowptr[r + 1] - beg);
   const int64_t* ord = rr->ord + beg;
   const double feastol = n->feastol;
   double fixedact = 0.0;
   double lhs, rhs;
   double minactinfvars = 0.0;
   double maxactinfvars = 0.0;
   long long gcd;
   int* infcheck;         /* positions v (sorted order) of the second group */
   int ninfcheckvars = 0;
   int nfixedconsvars = 0;
   int nunfixedvars;
   int ncontvars = 0;
   int gcdisone = 1;
   int possiblegcd = 1;
   int cutoff = 0;
   int v;

#define RR_COL(v_)   (p->colidx[ord[v_]])
#define RR_VAL(v_)   (p->vals[ord[v_]])
#define RR_FIXED(v_) (isEQ(n, lb[RR_COL(v_)], ub[RR_COL(v_)]))
#define RR_INT(v_)   (p->vartype[RR_COL(v_)] != 0)
#define RR_SECOND(v_) (!RR_INT(v_) || !epsIsInt(n, RR_VAL(v_)) || isEQ(n, fabs(RR_VAL(v_)), 1.0))

   if( !isRangedRow(p, n, r) )      /* :5771-5776 */
      return 0;

   /* fixed activity (:5805-5816), summed from the last nonzero to the first like the reference */
   for( v = nvars - 1; v >= 0; --v )
   {
      if( RR_FIXED(v) )
      {
         fixedact += lb[RR_COL(v)] * RR_VAL(v);
         ++nfixedconsvars;
      }
   }
   if( isHuge(n, fabs(fixedact)) )     /* :5819 */
      return 0;
   lhs = p->lhs[r] - fixedact;
   rhs = p->rhs[r] - fixedact;
   nunfixedvars = nvars - nfixedconsvars;
   infcheck = (int*)malloc(sizeof(int) * (size_t)(nvars + 1));

   /* partition (:5851-5893): everything in front of the first unfixed integer variable with an integral coefficient of
    * absolute value > 1 goes to the second group */
   v = -1;
   do
   {
      ++v;
      while( v < nvars && RR_SECOND(v) )
      {
         if( !RR_FIXED(v) )
         {
            if( !RR_INT(v) )
               ++ncontvars;
            gcdisone = gcdisone && isEQ(n, fabs(RR_VAL(v)), 1.0);
            possiblegcd = 0;
            infcheck[ninfcheckvars++] = v;
         }
         ++v;
      }
   }
   while( v < nvars && RR_FIXED(v) );

   if( v == nvars || ncontvars + 2 > nunfixedvars )      /* :5889, :5893 */
      goto TERMINATE;

   gcd = (long long)(fabs(RR_VAL(v)) + feastol);
   /* the rest (:5907-5957): gcd over the first group, what does not share a divisor joins the second */
   for( ; v < nvars; ++v )
   {
      if( RR_FIXED(v) )
         continue;
      if( RR_SECOND(v) )
      {
         if( !RR_INT(v) )
            ++ncontvars;
         gcdisone = gcdisone && isEQ(n, fabs(RR_VAL(v)), 1.0);
         possiblegcd = 0;
         infcheck[ninfcheckvars++] = v;
      }
      else
      {
         const long long gcdtmp = calcGcd(gcd, (long long)(fabs(RR_VAL(v)) + feastol));
         if( gcdtmp == 1 )
            infcheck[ninfcheckvars++] = v;
         else
            gcd = gcdtmp;
      }
   }
   if( ninfcheckvars == 0 )      /* :5962 */
      goto TERMINATE;

   /* activities of the second group (:5973-6018), last to first */
   for( v = ninfcheckvars - 1; v >= 0; --v )
   {
      const double l = lb[RR_COL(infcheck[v])];
      const double u = ub[RR_COL(infcheck[v])];
      const double a = RR_VAL(infcheck[v]);
      int mininvalid = 0;
      int maxinvalid = 0;
      if( isInf(n, -l) )
      {
         if( a < 0.0 ) maxinvalid = 1; else mininvalid = 1;
      }
      else
      {
         if( a < 0.0 ) maxactinfvars += a * l; else minactinfvars += a * l;
      }
      if( isInf(n, u) )
      {
         if( a > 0.0 ) maxinvalid = 1; else mininvalid = 1;
      }
      else
      {
         if( a > 0.0 ) maxactinfvars += a * u; else minactinfvars += a * u;
      }
      if( isHuge(n, -minactinfvars) )
         mininvalid = 1;
      if( isHuge(n, maxactinfvars) )
         maxinvalid = 1;
      if( mininvalid || maxinvalid )
         goto TERMINATE;
   }

   /* no multiple of the gcd between the sides (:6036-6047) */
   if( !epsIsInt(n, (lhs - maxactinfvars) / (double)gcd)
      && isGT(n, epsCeil(n, (lhs - maxactinfvars) / (double)gcd) * (double)gcd, rhs - minactinfvars) )
      cutoff = 1;
   else if( ncontvars == 0 )
   {
      long long gcdinfvars = -1;
      if( possiblegcd )
      {
         v = ninfcheckvars - 1;
         gcdinfvars = (long long)(fabs(RR_VAL(infcheck[v])) + feastol);
         for( ; v >= 0 && gcdinfvars >= 2; --v )
            gcdinfvars = calcGcd(gcdinfvars, (long long)(fabs(RR_VAL(infcheck[v])) + feastol));
      }
      else if( gcdisone )
         gcdinfvars = 1;

      if( gcdinfvars >= 1 )
      {
         double value;
         double value2;
         double minvalue = 0.0;
         double maxvalue = 0.0;
         int haveminvalue = 0;
         int nsols = 0;
         long long guard = 0;

         /* the values the second group can take, from below (:6072-6099) */
         value = epsCeil(n, minactinfvars - feastol);
         while( isLE(n, value, maxactinfvars) && ++guard < 100000000LL )
         {
            value2 = value + (double)gcd * epsCeil(n, (lhs - value) / (double)gcd);
            if( !isGE(n, value2, lhs) )
               value2 += (double)gcd;
            if( isLE(n, value2, rhs) )
            {
               ++nsols;
               if( nsols == 3 )
                  break;
               if( !haveminvalue )
               {
                  minvalue = value;
                  haveminvalue = 1;
               }
               maxvalue = value;
            }
            value += (double)gcdinfvars;
         }
         /* more than two: the last one from above (:6103-6130) */
         if( nsols == 3 )
         {
            guard = 0;
            value = epsFloor(n, maxactinfvars + feastol);
            while( isGE(n, value, minactinfvars) && ++guard < 100000000LL )
            {
               value2 = value + (double)gcd * epsFloor(n, (rhs - value) / (double)gcd);
               if( !isLE(n, value2, rhs) )
                  value2 -= (double)gcd;
               if( isGE(n, value2, lhs) )
               {
                  maxvalue = value;
                  break;
               }
               value -= (double)gcdinfvars;
            }
         }

         if( nsols == 0 )        /* :6136 */
            cutoff = 1;
         else
         {
            /* the single variable that can be bounded: the only one of the second group (:6156, :6323), or the only
             * unfixed one outside it (:6188, :6385) */
            int target = -1;
            int insecond = 0;
            if( ninfcheckvars == 1 )
            {
               target = infcheck[0];
               insecond = 1;
            }
            else if( ninfcheckvars == nunfixedvars - 1 )
            {
               int w = 0;
               for( v = 0; v < nvars; ++v )
               {
                  if( RR_FIXED(v) )
                     continue;
                  if( w < ninfcheckvars && infcheck[w] == v )
                  {
                     ++w;
                     continue;
                  }
                  target = v;
                  break;
               }
            }
            if( target >= 0 )
            {
               const int64_t j = RR_COL(target);
               const double a = RR_VAL(target);
               const int integral = p->vartype[j] != 0;
               double nlb, nub;
               if( nsols == 1 )
               {
                  /* SCIPinferVarFixCons (scip_var.c:6896): lower bound, then upper bound, both forced */
                  double fix;
                  if( insecond )
                     fix = maxvalue / a;                                                      /* :6172 */
                  else
                     fix = a < 0.0 ? epsFloor(n, (lhs - maxvalue) / a) : epsCeil(n, (lhs - maxvalue) / a);   /* :6224-6231 */
                  inferLb(n, integral, fix, lb[j], ub[j], 1, &newlb[j], &cutoff);
                  if( !cutoff )
                     inferUb(n, integral, fix, lb[j], ub[j], 1, &newub[j], &cutoff);
               }
               else
               {
                  if( insecond )
                  {
                     nlb = a < 0.0 ? maxvalue / a : minvalue / a;            /* :6332-6341 */
                     nub = a < 0.0 ? minvalue / a : maxvalue / a;
                  }
                  else if( a < 0.0 )
                  {
                     nlb = epsFloor(n, (rhs - minvalue) / a);                /* :6424-6427 */
                     nub = epsFloor(n, (lhs - maxvalue) / a);
                  }
                  else
                  {
                     nlb = epsCeil(n, (lhs - maxvalue) / a);                 /* :6431-6432 */
                     nub = epsCeil(n, (rhs - minvalue) / a);
                  }
                  if( nlb > lb[j] )
                     inferLb(n, integral, nlb, lb[j], ub[j], 1, &newlb[j], &cutoff);
                  if( !cutoff && nub < ub[j] )
                     inferUb(n, integral, nub, lb[j], ub[j], 1, &newub[j], &cutoff);
               }
            }
         }
      }
   }

 TERMINATE:
   free(infcheck);
   return cutoff;
#undef RR_COL
#undef RR_VAL
#undef RR_FIXED
#undef RR_INT
#undef RR_SECOND
}

static int sweepImpl(const ORACLE_PROBLEM* p, const ORACLE_NUMERICS* n, const RANGED* rr, const double* lb, const double* ub,
   double* newlb, double* newub, int64_t rowbegin, int64_t rowend);

int oracle_sweep(const ORACLE_PROBLEM* p, const ORACLE_NUMERICS* n, const double* lb, const double* ub,
   double* newlb, double* newub, int64_t rowbegin, int64_t rowend)
{
   return sweepImpl(p, n, NULL, lb, ub, newlb, newub, rowbegin, rowend);
}

static int sweepImpl(const ORACLE_PROBLEM* p, const ORACLE_NUMERICS* n, const RANGED* rr, const double* lb, const double* ub,
   double* newlb, double* newub, int64_t rowbegin, int64_t rowend)
{
   int64_t r;
   int cutoff = 0;

   for( r = rowbegin; r < rowend; ++r )
   {
      ROWACT ra;
      int64_t k;
      int64_t len = p->rowptr[r + 1] - p->rowptr[r];
      double lhs = p->lhs[r];
      double rhs = p->rhs[r];
      int tighten;
      int force = (len == 1);
      double minact, maxact;
      int t1, t2, s1, s2;

      rowActivities(p, n, lb, ub, r, &ra);

      /* tightenBounds gates: cons_linear.c:7021-7083 */
      tighten = !((ra.minneginf + ra.minposinf + ra.minneghuge + ra.minposhuge > 1)
         && (ra.maxneginf + ra.maxposinf + ra.maxneghuge + ra.maxposhuge > 1));
      if( tighten && isFeasZero(n, ra.maxactdelta) )
         tighten = 0;
      if( tighten && !isInf(n, ra.maxactdelta) )
      {
         double slack, surplus, m;
         getMinActivity(n, ra.minact, ra.minposinf, ra.minneginf, ra.minposhuge, ra.minneghuge, 0.0, 0, &minact, &t1, &s1);
         getMaxActivity(n, ra.maxact, ra.maxposinf, ra.maxneginf, ra.maxposhuge, ra.maxneghuge, 0.0, 0, &maxact, &t2, &s2);
         slack = (isInf(n, rhs) || s1) ? n->infinity : (rhs - minact);
         surplus = (isInf(n, -lhs) || s2) ? n->infinity : (maxact - lhs);
         m = slack < surplus ? slack : surplus;
         if( isLE(n, ra.maxactdelta, m) )
            tighten = 0;
      }

      if( tighten )
      {
         int easy = isLT(n, ra.maxactdelta, n->maxeasyactivitydelta);
         for( k = p->rowptr[r]; k < p->rowptr[r + 1] && !cutoff; ++k )
         {
            int32_t j = p->colidx[k];
            if( easy )
               cutoff = tightenVarBoundsEasy(n, &ra, p->vals[k], lhs, rhs, p->vartype[j] != 0, lb[j], ub[j], force, &newlb[j], &newub[j]);
            else
               cutoff = tightenVarBounds(n, &ra, p->vals[k], lhs, rhs, p->vartype[j] != 0, lb[j], ub[j], force, &newlb[j], &newub[j]);
         }
         if( cutoff )
            return 1;
      }

      /* ranged rows (propagateCons :7699-7712; tightenbounds is on) */
      if( rr != NULL && rangedRowPropagation(p, n, rr, lb, ub, r, newlb, newub) )
         return 1;

      /* row verdict: cons_linear.c:7715-7742 (goodrelax = TRUE) */
      getMinActivity(n, ra.minact, ra.minposinf, ra.minneginf, ra.minposhuge, ra.minneghuge, 0.0, 1, &minact, &t1, &s1);
      getMaxActivity(n, ra.maxact, ra.maxposinf, ra.maxneginf, ra.maxposhuge, ra.maxneghuge, 0.0, 1, &maxact, &t2, &s2);
      if( isFeasGT(n, minact, rhs) || isFeasLT(n, maxact, lhs) )
         return 1;
   }
   return 0;
}

/* the redundancy verdict of propagateCons, cons_linear.c:7715-7753: with the activities of the given bounds
 * (goodrelax = TRUE) a row that is not infeasible (FeasGT(minact, rhs) / FeasLT(maxact, lhs)) is redundant iff
 * GE(minact, lhs) and LE(maxact, rhs) -- the reference then deletes it locally (SCIPdelConsLocal, :7749) */
int64_t oracle_redundant_rows(const ORACLE_PROBLEM* p, const ORACLE_NUMERICS* n, const double* lb, const double* ub,
   uint8_t* redundant)
{
   int64_t r;
   int64_t count = 0;
   for( r = 0; r < p->nrows; ++r )
   {
      ROWACT ra;
      double minact;
      double maxact;
      int t1, t2, s1, s2;
      rowActivities(p, n, lb, ub, r, &ra);
      getMinActivity(n, ra.minact, ra.minposinf, ra.minneginf, ra.minposhuge, ra.minneghuge, 0.0, 1, &minact, &t1, &s1);
      getMaxActivity(n, ra.maxact, ra.maxposinf, ra.maxneginf, ra.maxposhuge, ra.maxneghuge, 0.0, 1, &maxact, &t2, &s2);
      redundant[r] = 0;
      if( isFeasGT(n, minact, p->rhs[r]) || isFeasLT(n, maxact, p->lhs[r]) )
         continue;
      if( isGE(n, minact, p->lhs[r]) && isLE(n, maxact, p->rhs[r]) )
      {
         redundant[r] = 1;
         ++count;
      }
   }
   return count;
}

static int propagateImpl(const ORACLE_PROBLEM* p, const ORACLE_NUMERICS* n, const RANGED* rr, double* lb, double* ub,
   int maxrounds, int* nrounds, int64_t* nchanges);

int oracle_propagate(const ORACLE_PROBLEM* p, const ORACLE_NUMERICS* n, double* lb, double* ub, int maxrounds,
   int* nrounds, int64_t* nchanges)
{
   return propagateImpl(p, n, NULL, lb, ub, maxrounds, nrounds, nchanges);
}

/* the sorted order of every ranged row (see RANGED); sortlb / sortub: the global bounds the reference sorts by, tie: the
 * SCIPvarGetProbindex of every column or NULL (the column index) */
int64_t oracle_ranged_order(const ORACLE_PROBLEM* p, const ORACLE_NUMERICS* n, const double* sortlb, const double* sortub,
   const int32_t* tie, int64_t* ord)
{
   int64_t r;
   int64_t k;
   int64_t nranged = 0;
   for( k = 0; k < p->nnz; ++k )
      ord[k] = k;
   g_sortprob = p;
   g_sortglb = sortlb;
   g_sortgub = sortub;
   g_sorttie = tie;
   for( r = 0; r < p->nrows; ++r )
   {
      if( isRangedRow(p, n, r) )
      {
         qsort(ord + p->rowptr[r], (size_t)(p->rowptr[r + 1] - p->rowptr[r]), sizeof(int64_t), compVarProp);
         ++nranged;
      }
   }
   return nranged;
}

int oracle_propagate_ranged(const ORACLE_PROBLEM* p, const ORACLE_NUMERICS* n, double* lb, double* ub, int maxrounds,
   int* nrounds, int64_t* nchanges, const double* sortlb, const double* sortub, const int32_t* tie)
{
   RANGED rr;
   int64_t* ord = (int64_t*)malloc(sizeof(int64_t) * (size_t)(p->nnz + 1));
   int status;
   oracle_ranged_order(p, n, sortlb != NULL ? sortlb : lb, sortub != NULL ? sortub : ub, tie, ord);
   rr.ord = ord;
   status = propagateImpl(p, n, &rr, lb, ub, maxrounds, nrounds, nchanges);
   free(ord);
   return status;
}

static int propagateImpl(const ORACLE_PROBLEM* p, const ORACLE_NUMERICS* n, const RANGED* rr, double* lb, double* ub,
   int maxrounds, int* nrounds, int64_t* nchanges)
{
   double* newlb = (double*)malloc(sizeof(double) * (size_t)(p->ncols + 1));
   double* newub = (double*)malloc(sizeof(double) * (size_t)(p->ncols + 1));
   int status = ORACLE_STATUS_ROUNDLIMIT;
   int round = 0;
   int64_t total = 0;
   int64_t j;

   for( j = 0; j < p->ncols; ++j )
   {
      /* canonicalise -0.0 */
      lb[j] += 0.0;
      ub[j] += 0.0;
   }

   while( maxrounds <= 0 || round < maxrounds )
   {
      int64_t nchg = 0;
      int cutoff;

      memcpy(newlb, lb, sizeof(double) * (size_t)p->ncols);
      memcpy(newub, ub, sizeof(double) * (size_t)p->ncols);
      ++round;
      cutoff = sweepImpl(p, n, rr, lb, ub, newlb, newub, 0, p->nrows);
      if( cutoff )
      {
         status = ORACLE_STATUS_CUTOFF;
         break;
      }
      /* apply: both bounds of a variable may have moved in the same round */
      for( j = 0; j < p->ncols && !cutoff; ++j )
      {
         if( newlb[j] > newub[j] )
         {
            if( isFeasGT(n, newlb[j], newub[j]) )
               cutoff = 1;
            else
               newlb[j] = newub[j];
         }
         if( newlb[j] != lb[j] )
            ++nchg;
         if( newub[j] != ub[j] )
            ++nchg;
      }
      if( cutoff )
      {
         status = ORACLE_STATUS_CUTOFF;
         break;
      }
      memcpy(lb, newlb, sizeof(double) * (size_t)p->ncols);
      memcpy(ub, newub, sizeof(double) * (size_t)p->ncols);
      total += nchg;
      if( nchg == 0 )
      {
         status = ORACLE_STATUS_FIXPOINT;
         break;
      }
   }
   free(newlb);
   free(newub);
   if( nrounds != NULL )
      *nrounds = round;
   if( nchanges != NULL )
      *nchanges = total;
   return status;
}
