/* linprop_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C, single thread) of the reference's activity-based bound propagation over linear
 * constraints (cons_linear.c tightenBounds/activity path), run as synchronous (Jacobi) rounds to a fixpoint.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/--impl reference legs may use it; the
 * product (scip_b200/) never includes, links or loads anything under oracle/.
 *
 * Parity status: PINNED -- checked against the compiled reference itself (oracle/_ref, built by
 * oracle/Makefile.ref from /root/reference) on every check/instances/MIP/ *.mps plus seeded synthetic
 * instances; fixtures under tests/golden/ (see tests/golden/make_golden.py).
 */
#ifndef LINPROP_ORACLE_H
#define LINPROP_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/** numerical tolerances; defaults are the reference's (def.h:172-184, cons_linear.c:135) */
typedef struct
{
   double infinity;             /* 1e20  */
   double epsilon;              /* 1e-9  */
   double sumepsilon;           /* 1e-6  */
   double feastol;              /* 1e-6  */
   double boundstreps;          /* 0.05  */
   double hugeval;              /* 1e15  */
   double maxeasyactivitydelta; /* 1e6   */
} ORACLE_NUMERICS;

typedef struct
{
   int64_t        nrows;
   int64_t        ncols;
   int64_t        nnz;
   const int64_t* rowptr;   /* nrows+1 */
   const int32_t* colidx;   /* nnz */
   const double*  vals;     /* nnz */
   const double*  lhs;      /* nrows, -infinity for none */
   const double*  rhs;      /* nrows, +infinity for none */
   const uint8_t* vartype;  /* ncols: 0 continuous, 1 integral (SCIPvarIsIntegral) */
} ORACLE_PROBLEM;

#define ORACLE_STATUS_FIXPOINT   0
#define ORACLE_STATUS_CUTOFF     1
#define ORACLE_STATUS_ROUNDLIMIT 2

void oracle_default_numerics(ORACLE_NUMERICS* num);

/** propagates lb/ub (in place) to the Jacobi fixpoint; returns ORACLE_STATUS_* */
int oracle_propagate(
   const ORACLE_PROBLEM*  prob,
   const ORACLE_NUMERICS* num,
   double*                lb,         /* ncols, in/out */
   double*                ub,         /* ncols, in/out */
   int                    maxrounds,  /* <= 0: unlimited */
   int*                   nrounds,    /* out: sweeps executed (including the final one without change) */
   int64_t*               nchanges    /* out: number of accepted bound changes (per variable and round) */
   );

/** the same with ranged-row propagation switched on (rangedRowPropagation, cons_linear.c:5715-6696, with
 *  constraints/linear/rangedrowartcons = FALSE): after the bound tightening of a row with two finite sides and at least
 *  three nonzeros the gcd rule runs on it.  The rule walks the row in the order the reference has sorted it in
 *  (consdataCompVarProp, cons_linear.c:3191): sortlb / sortub = the global bounds at that time (NULL: lb / ub on entry),
 *  tie = SCIPvarGetProbindex of every column (NULL: the column index) */
int oracle_propagate_ranged(
   const ORACLE_PROBLEM*  prob,
   const ORACLE_NUMERICS* num,
   double*                lb,
   double*                ub,
   int                    maxrounds,
   int*                   nrounds,
   int64_t*               nchanges,
   const double*          sortlb,
   const double*          sortub,
   const int32_t*         tie
   );

/** the order itself: ord[rowptr[r] + v] = position of the v-th nonzero of row r in the reference's sorted order (identity
 *  for rows the ranged-row rule does not look at); returns the number of ranged rows */
int64_t oracle_ranged_order(
   const ORACLE_PROBLEM*  prob,
   const ORACLE_NUMERICS* num,
   const double*          sortlb,
   const double*          sortub,
   const int32_t*         tie,
   int64_t*               ord        /* nnz, out */
   );

/** one synchronous sweep over rows [rowbegin,rowend): reads lb/ub, writes improved bounds into newlb/newub
 *  (which must hold copies of lb/ub on entry); returns 1 if a cutoff was detected, else 0 */
int oracle_sweep(
   const ORACLE_PROBLEM*  prob,
   const ORACLE_NUMERICS* num,
   const double*          lb,
   const double*          ub,
   double*                newlb,
   double*                newub,
   int64_t                rowbegin,
   int64_t                rowend
   );

/** redundant[r] = 1 iff row r is redundant for the bounds lb/ub in the sense of propagateCons (cons_linear.c:7743:
 *  not infeasible, GE(minactivity, lhs) and LE(maxactivity, rhs)); returns their number */
int64_t oracle_redundant_rows(
   const ORACLE_PROBLEM*  prob,
   const ORACLE_NUMERICS* num,
   const double*          lb,
   const double*          ub,
   uint8_t*               redundant   /* nrows, out */
   );

/** double-double helpers exported for the known-answer tests (dbldblarith.h:154-187) */
void oracle_dd_sum21(double* rhi, double* rlo, double ahi, double alo, double b);

#ifdef __cplusplus
}
#endif

#endif
