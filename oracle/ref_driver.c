/* ref_driver.c -- TEST/BENCH INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Runs the UNMODIFIED reference (libscip built by oracle/Makefile.ref) on one instance with the
 * parity settings of SURVEY.md section 8c: root-node cons_linear propagation to fixpoint, presolve/LP/heuristics/
 * other propagators off.  It
 *   (1) loads an instance: a file any reference reader understands (--read X.mps) or a ".lpb" problem
 *       file (--lpb X.lpb, see oracle/lpb_format.md) that is rebuilt through SCIPcreateVarBasic /
 *       SCIPcreateConsBasicLinear,
 *   (2) optionally dumps the linear rows exactly as a propagator plugin sees them at the first root
 *       propagation call (--dump-lpb OUT.lpb) -- SCIPgetVarsLinear/SCIPgetValsLinear/SCIPgetLhsLinear/
 *       SCIPgetRhsLinear (cons_linear.h:263-307), SCIPvarGetLbLocal/UbLocal, SCIPvarIsIntegral,
 *   (3) lets the reference propagate (propagateDomains, solve.c:723; consPropLinear, cons_linear.c:16126),
 *   (4) writes the resulting global bounds + verdict + SCIPconshdlrGetPropTime (--out OUT.lpr).
 * Variables are indexed by their position in SCIPgetOrigVars in every dump.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <time.h>

#include "scip/scip.h"
#include "scip/scipdefplugins.h"

#define LPB_MAGIC "GPULPB01"
#define LPR_MAGIC "GPULPR01"

typedef struct
{
   int64_t nrows, ncols, nnz;
   int64_t* rowptr;
   int32_t* colidx;
   double*  vals;
   double*  lhs;
   double*  rhs;
   double*  lb;
   double*  ub;
   uint8_t* vartype;   /* 0 continuous, 1 integral */
} LPB;

struct SCIP_PropData
{
   const char* dumpfile;
   SCIP_VAR**  origvars;
   int         norigvars;
   int         dumped;
};

static double wallclock(void)
{
   struct timespec ts;
   clock_gettime(CLOCK_MONOTONIC, &ts);
   return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static int lpb_write(const char* fn, const LPB* p)
{
   FILE* f = fopen(fn, "wb");
   size_t pad;
   static const char zeros[8] = {0};
   if( f == NULL )
      return 0;
   fwrite(LPB_MAGIC, 1, 8, f);
   fwrite(&p->nrows, 8, 1, f);
   fwrite(&p->ncols, 8, 1, f);
   fwrite(&p->nnz, 8, 1, f);
   fwrite(p->rowptr, 8, (size_t)p->nrows + 1, f);
   fwrite(p->colidx, 4, (size_t)p->nnz, f);
   pad = (size_t)((8 - (4 * p->nnz) % 8) % 8);
   fwrite(zeros, 1, pad, f);
   fwrite(p->vals, 8, (size_t)p->nnz, f);
   fwrite(p->lhs, 8, (size_t)p->nrows, f);
   fwrite(p->rhs, 8, (size_t)p->nrows, f);
   fwrite(p->lb, 8, (size_t)p->ncols, f);
   fwrite(p->ub, 8, (size_t)p->ncols, f);
   fwrite(p->vartype, 1, (size_t)p->ncols, f);
   fclose(f);
   return 1;
}

static int lpb_read(const char* fn, LPB* p)
{
   FILE* f = fopen(fn, "rb");
   char magic[8];
   size_t pad;
   char skip[8];
   if( f == NULL )
      return 0;
   if( fread(magic, 1, 8, f) != 8 || memcmp(magic, LPB_MAGIC, 8) != 0 )
   {
      fclose(f);
      return 0;
   }
   if( fread(&p->nrows, 8, 1, f) != 1 || fread(&p->ncols, 8, 1, f) != 1 || fread(&p->nnz, 8, 1, f) != 1 )
   {
      fclose(f);
      return 0;
   }
   p->rowptr = (int64_t*)malloc(8 * ((size_t)p->nrows + 1));
   p->colidx = (int32_t*)malloc(4 * (size_t)p->nnz + 8);
   p->vals = (double*)malloc(8 * (size_t)p->nnz + 8);
   p->lhs = (double*)malloc(8 * (size_t)p->nrows + 8);
   p->rhs = (double*)malloc(8 * (size_t)p->nrows + 8);
   p->lb = (double*)malloc(8 * (size_t)p->ncols + 8);
   p->ub = (double*)malloc(8 * (size_t)p->ncols + 8);
   p->vartype = (uint8_t*)malloc((size_t)p->ncols + 8);
   pad = (size_t)((8 - (4 * p->nnz) % 8) % 8);
   if( fread(p->rowptr, 8, (size_t)p->nrows + 1, f) != (size_t)p->nrows + 1
      || fread(p->colidx, 4, (size_t)p->nnz, f) != (size_t)p->nnz
      || fread(skip, 1, pad, f) != pad
      || fread(p->vals, 8, (size_t)p->nnz, f) != (size_t)p->nnz
      || fread(p->lhs, 8, (size_t)p->nrows, f) != (size_t)p->nrows
      || fread(p->rhs, 8, (size_t)p->nrows, f) != (size_t)p->nrows
      || fread(p->lb, 8, (size_t)p->ncols, f) != (size_t)p->ncols
      || fread(p->ub, 8, (size_t)p->ncols, f) != (size_t)p->ncols
      || fread(p->vartype, 1, (size_t)p->ncols, f) != (size_t)p->ncols )
   {
      fclose(f);
      return 0;
   }
   fclose(f);
   return 1;
}

/** index of a transformed variable in original-variable order (stored in a side table keyed by probindex) */
static int* g_probidx2orig = NULL;

/* --dump-tie FILE: SCIPvarGetProbindex of every variable (.lpb order) at the first root propagation -- the last key by
 * which the reference sorts the nonzeros of a row (consdataCompVarProp, cons_linear.c:3191), which the ranged-row rule walks */
static const char* g_tiefile = NULL;
static int32_t* g_tie = NULL;

/** dump the linear rows as seen at the first propagation call */
static SCIP_RETCODE dumpProblem(SCIP* scip, SCIP_PROPDATA* propdata)
{
   SCIP_CONSHDLR* conshdlr;
   SCIP_CONS** conss;
   LPB p;
   int nconss;
   int ntransvars;
   int c;
   int i;
   int64_t k;

   ntransvars = SCIPgetNVars(scip);
   g_probidx2orig = (int*)malloc(sizeof(int) * (size_t)(ntransvars + 1));
   if( g_tiefile != NULL )
      g_tie = (int32_t*)malloc(sizeof(int32_t) * (size_t)(propdata->norigvars + 1));
   for( i = 0; i < ntransvars; ++i )
      g_probidx2orig[i] = -1;

   p.ncols = propdata->norigvars;
   p.lb = (double*)malloc(8 * (size_t)p.ncols + 8);
   p.ub = (double*)malloc(8 * (size_t)p.ncols + 8);
   p.vartype = (uint8_t*)malloc((size_t)p.ncols + 8);
   for( i = 0; i < propdata->norigvars; ++i )
   {
      SCIP_VAR* tv = SCIPvarGetTransVar(propdata->origvars[i]);
      int pi;
      if( tv == NULL || (pi = SCIPvarGetProbindex(tv)) < 0 )
      {
         fprintf(stderr, "ref_driver: original variable %d has no active transformed counterpart\n", i);
         return SCIP_ERROR;
      }
      g_probidx2orig[pi] = i;
      if( g_tiefile != NULL )
         g_tie[i] = pi;
      p.lb[i] = SCIPvarGetLbLocal(tv);
      p.ub[i] = SCIPvarGetUbLocal(tv);
      p.vartype[i] = SCIPvarIsIntegral(tv) ? 1 : 0;
   }

   conshdlr = SCIPfindConshdlr(scip, "linear");
   conss = SCIPconshdlrGetConss(conshdlr);
   nconss = SCIPconshdlrGetNActiveConss(conshdlr);
   p.nrows = nconss;
   p.nnz = 0;
   for( c = 0; c < nconss; ++c )
      p.nnz += SCIPgetNVarsLinear(scip, conss[c]);
   p.rowptr = (int64_t*)malloc(8 * ((size_t)p.nrows + 1));
   p.colidx = (int32_t*)malloc(4 * (size_t)p.nnz + 8);
   p.vals = (double*)malloc(8 * (size_t)p.nnz + 8);
   p.lhs = (double*)malloc(8 * (size_t)p.nrows + 8);
   p.rhs = (double*)malloc(8 * (size_t)p.nrows + 8);
   k = 0;
   for( c = 0; c < nconss; ++c )
   {
      SCIP_VAR** vars = SCIPgetVarsLinear(scip, conss[c]);
      SCIP_Real* vals = SCIPgetValsLinear(scip, conss[c]);
      int nv = SCIPgetNVarsLinear(scip, conss[c]);
      int v;
      if( SCIPconsIsModifiable(conss[c]) || SCIPconsIsLocal(conss[c]) )
      {
         fprintf(stderr, "ref_driver: modifiable/local linear row %d is not supported by the dump\n", c);
         return SCIP_ERROR;
      }
      p.rowptr[c] = k;
      p.lhs[c] = SCIPgetLhsLinear(scip, conss[c]);
      p.rhs[c] = SCIPgetRhsLinear(scip, conss[c]);
      for( v = 0; v < nv; ++v )
      {
         int pi = SCIPvarGetProbindex(vars[v]);
         if( pi < 0 || g_probidx2orig[pi] < 0 )
         {
            fprintf(stderr, "ref_driver: row %d holds a non-active variable\n", c);
            return SCIP_ERROR;
         }
         p.colidx[k] = g_probidx2orig[pi];
         p.vals[k] = vals[v];
         ++k;
      }
   }
   p.rowptr[nconss] = k;
   if( propdata->dumpfile != NULL && !lpb_write(propdata->dumpfile, &p) )
   {
      fprintf(stderr, "ref_driver: cannot write %s\n", propdata->dumpfile);
      return SCIP_ERROR;
   }
   free(p.rowptr); free(p.colidx); free(p.vals); free(p.lhs); free(p.rhs); free(p.lb); free(p.ub); free(p.vartype);
   return SCIP_OKAY;
}

static SCIP_DECL_PROPEXEC(propExecDump)
{
   SCIP_PROPDATA* propdata = SCIPpropGetData(prop);
   (void)proptiming;
   *result = SCIP_DIDNOTRUN;
   if( !propdata->dumped && SCIPgetDepth(scip) == 0 )
   {
      propdata->dumped = 1;
      SCIP_CALL( dumpProblem(scip, propdata) );
   }
   return SCIP_OKAY;
}

/** variables in .lpb (creation) order; SCIPgetOrigVars is sorted by variable type instead */
static SCIP_VAR** g_lpbvars = NULL;
/* --reprop (see propExecReprop) */
static int g_reprop = 0;
static int g_repropdone = 0;
static int g_nfix = 0;
static int* g_fixvar = NULL;      /* index into g_lpbvars */
static double* g_fixval = NULL;

static SCIP_RETCODE buildFromLpb(SCIP* scip, const LPB* p)
{
   SCIP_VAR** vars;
   int64_t i;
   char name[64];

   SCIP_CALL( SCIPcreateProbBasic(scip, "lpb") );
   vars = (SCIP_VAR**)malloc(sizeof(SCIP_VAR*) * (size_t)p->ncols);
   for( i = 0; i < p->ncols; ++i )
   {
      SCIP_VARTYPE vt = SCIP_VARTYPE_CONTINUOUS;
      double lb = p->lb[i];
      double ub = p->ub[i];
      if( p->vartype[i] )
         vt = (lb == 0.0 && ub == 1.0) ? SCIP_VARTYPE_BINARY : SCIP_VARTYPE_INTEGER;
      if( g_reprop > 0 && p->vartype[i] && lb == ub && (lb == 0.0 || lb == 1.0) )
      {
         /* --reprop: a fixed binary is created free; its fixing is applied as a probing bound change in every cycle */
         if( g_fixvar == NULL )
         {
            g_fixvar = (int*)malloc(sizeof(int) * (size_t)(p->ncols + 1));
            g_fixval = (double*)malloc(sizeof(double) * (size_t)(p->ncols + 1));
         }
         g_fixvar[g_nfix] = (int)i;
         g_fixval[g_nfix++] = lb;
         lb = 0.0;
         ub = 1.0;
         vt = SCIP_VARTYPE_BINARY;
      }
      snprintf(name, sizeof(name), "x%lld", (long long)i);
      SCIP_CALL( SCIPcreateVarBasic(scip, &vars[i], name, lb, ub, 0.0, vt) );
      SCIP_CALL( SCIPaddVar(scip, vars[i]) );
   }
   {
      int maxlen = 0;
      SCIP_VAR** rowvars;
      for( i = 0; i < p->nrows; ++i )
         if( p->rowptr[i + 1] - p->rowptr[i] > maxlen )
            maxlen = (int)(p->rowptr[i + 1] - p->rowptr[i]);
      rowvars = (SCIP_VAR**)malloc(sizeof(SCIP_VAR*) * (size_t)(maxlen + 1));
      for( i = 0; i < p->nrows; ++i )
      {
         SCIP_CONS* cons;
         int len = (int)(p->rowptr[i + 1] - p->rowptr[i]);
         int v;
         for( v = 0; v < len; ++v )
            rowvars[v] = vars[p->colidx[p->rowptr[i] + v]];
         snprintf(name, sizeof(name), "c%lld", (long long)i);
         SCIP_CALL( SCIPcreateConsBasicLinear(scip, &cons, name, len, rowvars, p->vals + p->rowptr[i], p->lhs[i], p->rhs[i]) );
         SCIP_CALL( SCIPaddCons(scip, cons) );
         SCIP_CALL( SCIPreleaseCons(scip, &cons) );
      }
      free(rowvars);
   }
   g_lpbvars = (SCIP_VAR**)malloc(sizeof(SCIP_VAR*) * (size_t)(p->ncols + 1));
   memcpy(g_lpbvars, vars, sizeof(SCIP_VAR*) * (size_t)p->ncols);
   for( i = 0; i < p->ncols; ++i )
      SCIP_CALL( SCIPreleaseVar(scip, &vars[i]) );
   free(vars);
   return SCIP_OKAY;
}

/** the parity settings of SURVEY.md 8c -- all ordinary reference parameters */
static SCIP_RETCODE applyParitySettings(SCIP* scip, double boundstreps, int quiet)
{
   static const char* offprops[] = { "dualfix", "genvbounds", "nlobbt", "obbt", "probing", "pseudoobj", "redcost",
      "rootredcost", "vbounds", "symmetry", NULL };
   char pname[128];
   int i;

   SCIP_CALL( SCIPsetIntParam(scip, "presolving/maxrounds", 0) );
   SCIP_CALL( SCIPsetIntParam(scip, "presolving/maxrestarts", 0) );
   SCIP_CALL( SCIPsetIntParam(scip, "propagating/maxrounds", -1) );
   SCIP_CALL( SCIPsetIntParam(scip, "propagating/maxroundsroot", -1) );
   SCIP_CALL( SCIPsetIntParam(scip, "lp/solvefreq", -1) );
   SCIP_CALL( SCIPsetLongintParam(scip, "limits/nodes", 1LL) );
   SCIP_CALL( SCIPsetBoolParam(scip, "conflict/enable", FALSE) );
   SCIP_CALL( SCIPsetBoolParam(scip, "constraints/linear/rangedrowpropagation", FALSE) );
   SCIP_CALL( SCIPsetIntParam(scip, "timing/clocktype", 2) );
   if( SCIPgetParam(scip, "misc/usesymmetry") != NULL )
      SCIP_CALL( SCIPsetIntParam(scip, "misc/usesymmetry", 0) );
   for( i = 0; offprops[i] != NULL; ++i )
   {
      snprintf(pname, sizeof(pname), "propagating/%s/freq", offprops[i]);
      if( SCIPgetParam(scip, pname) != NULL )
         SCIP_CALL( SCIPsetIntParam(scip, pname, -1) );
   }
   SCIP_CALL( SCIPsetHeuristics(scip, SCIP_PARAMSETTING_OFF, TRUE) );
   SCIP_CALL( SCIPsetSeparating(scip, SCIP_PARAMSETTING_OFF, TRUE) );
   if( boundstreps > 0.0 )
      SCIP_CALL( SCIPsetRealParam(scip, "numerics/boundstreps", boundstreps) );
   if( quiet )
      SCIP_CALL( SCIPsetIntParam(scip, "display/verblevel", 0) );
   return SCIP_OKAY;
}

/* --probe N: once the root node is propagated (a delayed propagator), the reference's own probing cycle for up to N
 * unfixed integral variables -- SCIPstartProbing / SCIPchgVarLb|UbProbing / SCIPpropagateProbing / SCIPendProbing, what
 * SCIPapplyProbingVar does per candidate (prop_probing.c:1254-1279) -- timed: the CPU baseline of BASELINE configs[4] */
static int g_nprobe = 0;
static int g_probedone = 0;
static int g_inprobe = 0;

static SCIP_DECL_PROPEXEC(propExecProbe)
{
   SCIP_VAR** vars = SCIPgetVars(scip);
   int nvars = SCIPgetNVars(scip);
   int n = 0;
   int ncut = 0;
   SCIP_Longint ndomtotal = 0;
   double t0, t1;
   int i;

   (void)prop;
   (void)proptiming;
   *result = SCIP_DIDNOTRUN;
   if( g_probedone || g_inprobe || SCIPgetDepth(scip) != 0 || SCIPinProbing(scip) )
      return SCIP_OKAY;
   g_probedone = 1;
   g_inprobe = 1;
   t0 = wallclock();
   /* every (nvars / N)-th candidate, so that the probes are spread over the columns */
   for( i = 0; i < nvars && n < g_nprobe; i += (nvars > 8 * g_nprobe ? 7 : 1) )
   {
      const SCIP_Real lb = SCIPvarGetLbLocal(vars[i]);
      const SCIP_Real ub = SCIPvarGetUbLocal(vars[i]);
      SCIP_Bool cutoff;
      SCIP_Longint ndom;
      if( !SCIPvarIsIntegral(vars[i]) || ub - lb < 0.5 || SCIPisInfinity(scip, -lb) || SCIPisInfinity(scip, ub) )
         continue;
      SCIP_CALL( SCIPstartProbing(scip) );
      if( n % 2 == 0 )
         SCIP_CALL( SCIPchgVarLbProbing(scip, vars[i], SCIPfeasFloor(scip, 0.5 * (lb + ub)) + 1.0) );
      else
         SCIP_CALL( SCIPchgVarUbProbing(scip, vars[i], SCIPfeasFloor(scip, 0.5 * (lb + ub))) );
      SCIP_CALL( SCIPpropagateProbing(scip, -1, &cutoff, &ndom) );
      SCIP_CALL( SCIPendProbing(scip) );
      ncut += cutoff ? 1 : 0;
      ndomtotal += ndom;
      ++n;
   }
   t1 = wallclock();
   printf("PROBING {\"probes\": %d, \"cutoffs\": %d, \"domreds\": %lld, \"probing_s\": %.9g}\n", n, ncut,
      (long long)ndomtotal, t1 - t0);
   g_inprobe = 0;
   *result = SCIP_DIDNOTFIND;
   return SCIP_OKAY;
}

/* --reprop K (with --lpb): the root propagation of the instance, K times inside ONE solve.  Transforming and freeing a
 * model of 10M nonzeros costs the reference 30 s per solve around 6 s of propagation; here the binaries the instance fixes
 * (lb == ub) are created free, the root node is propagated (nothing to find), and then K times: SCIPstartProbing, the
 * instance's fixings as probing bound changes, SCIPpropagateProbing to the fixpoint -- propagationRound (solve.c:437) ->
 * consPropLinear -> propagateCons -> tightenBounds, the same path on the same rows and bounds -- SCIPendProbing.  One
 * REPROP line per cycle: SCIPconshdlrGetPropTime(linear) spent in it, wall time of SCIPpropagateProbing, reductions. */
static double* g_reproplb = NULL; /* local bounds at the fixpoint of the last cycle, .lpb order */
static double* g_repropub = NULL;
static int g_repropcutoff = 0;
static SCIP_Longint g_repropdomreds = 0;
static double g_reproptime = 0.0;

static SCIP_DECL_PROPEXEC(propExecReprop)
{
   SCIP_CONSHDLR* linhdlr = SCIPfindConshdlr(scip, "linear");
   int norig = SCIPgetNOrigVars(scip);
   int rep;
   int i;

   (void)prop;
   (void)proptiming;
   *result = SCIP_DIDNOTRUN;
   if( g_repropdone || SCIPgetDepth(scip) != 0 || SCIPinProbing(scip) )
      return SCIP_OKAY;
   g_repropdone = 1;
   g_reproplb = (double*)malloc(8 * (size_t)norig + 8);
   g_repropub = (double*)malloc(8 * (size_t)norig + 8);
   for( rep = 0; rep < g_reprop; ++rep )
   {
      SCIP_Bool cutoff = FALSE;
      SCIP_Longint ndom = 0;
      const double p0 = SCIPconshdlrGetPropTime(linhdlr);
      const SCIP_Longint c0 = SCIPconshdlrGetNPropCalls(linhdlr);
      double t0, t1, t2;

      t0 = wallclock();
      SCIP_CALL( SCIPstartProbing(scip) );
      for( i = 0; i < g_nfix; ++i )
      {
         SCIP_VAR* tv = SCIPvarGetTransVar(g_lpbvars[g_fixvar[i]]);
         if( g_fixval[i] < 0.5 )
            SCIP_CALL( SCIPchgVarUbProbing(scip, tv, g_fixval[i]) );
         else
            SCIP_CALL( SCIPchgVarLbProbing(scip, tv, g_fixval[i]) );
      }
      t1 = wallclock();
      SCIP_CALL( SCIPpropagateProbing(scip, -1, &cutoff, &ndom) );
      t2 = wallclock();
      g_reproptime = SCIPconshdlrGetPropTime(linhdlr) - p0;
      g_repropcutoff = cutoff ? 1 : 0;
      g_repropdomreds = ndom;
      if( rep + 1 == g_reprop )
      {
         for( i = 0; i < norig; ++i )
         {
            SCIP_VAR* tv = SCIPvarGetTransVar(g_lpbvars[i]);
            g_reproplb[i] = SCIPvarGetLbLocal(tv);
            g_repropub[i] = SCIPvarGetUbLocal(tv);
         }
      }
      SCIP_CALL( SCIPendProbing(scip) );
      printf("REPROP {\"prop_time_s\": %.9g, \"propagate_wall_s\": %.9g, \"fixings_wall_s\": %.9g, \"backtrack_wall_s\": %.9g, "
         "\"prop_calls\": %lld, \"domreds\": %lld, \"cutoff\": %d, \"fixings\": %d}\n", g_reproptime, t2 - t1, t1 - t0,
         wallclock() - t2, (long long)(SCIPconshdlrGetNPropCalls(linhdlr) - c0), (long long)ndom, g_repropcutoff, g_nfix);
      fflush(stdout);
   }
   *result = SCIP_DIDNOTFIND;
   return SCIP_OKAY;
}

static SCIP_RETCODE run(int argc, char** argv)
{
   SCIP* scip = NULL;
   SCIP_PROP* prop = NULL;
   SCIP_PROPDATA propdata;
   SCIP_CONSHDLR* linhdlr;
   const char* readfile = NULL;
   const char* lpbfile = NULL;
   const char* outfile = NULL;
   const char* dumpfile = NULL;
   const char* redfile = NULL;
   double boundstreps = -1.0;
   int quiet = 1;
   int repeat = 1;
   int rangedrow = 0;
   int rep;
   int i;
   double t0, t1, tbuild;
   double ttrans = 0.0, tpre = 0.0;
   int infeasible;
   FILE* f;

   for( i = 1; i < argc; ++i )
   {
      if( strcmp(argv[i], "--read") == 0 && i + 1 < argc ) readfile = argv[++i];
      else if( strcmp(argv[i], "--lpb") == 0 && i + 1 < argc ) lpbfile = argv[++i];
      else if( strcmp(argv[i], "--out") == 0 && i + 1 < argc ) outfile = argv[++i];
      else if( strcmp(argv[i], "--dump-lpb") == 0 && i + 1 < argc ) dumpfile = argv[++i];
      else if( strcmp(argv[i], "--boundstreps") == 0 && i + 1 < argc ) boundstreps = atof(argv[++i]);
      else if( strcmp(argv[i], "--verbose") == 0 ) quiet = 0;
      else if( strcmp(argv[i], "--probe") == 0 && i + 1 < argc ) g_nprobe = atoi(argv[++i]);
      else if( strcmp(argv[i], "--redundant") == 0 && i + 1 < argc ) redfile = argv[++i];
      else if( strcmp(argv[i], "--repeat") == 0 && i + 1 < argc ) repeat = atoi(argv[++i]);
      else if( strcmp(argv[i], "--rangedrow") == 0 ) rangedrow = 1;
      else if( strcmp(argv[i], "--dump-tie") == 0 && i + 1 < argc ) g_tiefile = argv[++i];
      else if( strcmp(argv[i], "--reprop") == 0 && i + 1 < argc ) g_reprop = atoi(argv[++i]);
      else
      {
         fprintf(stderr, "usage: ref_driver (--read FILE | --lpb FILE) [--out OUT.lpr] [--dump-lpb OUT.lpb] [--boundstreps X] [--repeat K] [--verbose]\n");
         return SCIP_ERROR;
      }
   }
   if( (readfile == NULL) == (lpbfile == NULL) )
   {
      fprintf(stderr, "ref_driver: give exactly one of --read / --lpb\n");
      return SCIP_ERROR;
   }

   SCIP_CALL( SCIPcreate(&scip) );
   SCIP_CALL( SCIPincludeDefaultPlugins(scip) );
   if( g_nprobe > 0 )
   {
      SCIP_PROP* probeprop = NULL;
      SCIP_CALL( SCIPincludePropBasic(scip, &probeprop, "refprobe", "times the reference's probing cycle at the propagated root",
            -1000, 1, TRUE, SCIP_PROPTIMING_BEFORELP, propExecProbe, NULL) );
   }
   if( g_reprop > 0 )
   {
      SCIP_PROP* repprop = NULL;
      if( lpbfile == NULL || repeat > 1 )
      {
         fprintf(stderr, "ref_driver: --reprop needs --lpb and excludes --repeat\n");
         return SCIP_ERROR;
      }
      SCIP_CALL( SCIPincludePropBasic(scip, &repprop, "refreprop", "repeats the root propagation of the instance in probing mode",
            -1000, 1, TRUE, SCIP_PROPTIMING_BEFORELP, propExecReprop, NULL) );
   }
   memset(&propdata, 0, sizeof(propdata));
   propdata.dumpfile = dumpfile;
   SCIP_CALL( SCIPincludePropBasic(scip, &prop, "refdump", "dumps linear rows at the first root propagation call",
         100000000, 1, FALSE, SCIP_PROPTIMING_BEFORELP, propExecDump, &propdata) );
   SCIP_CALL( applyParitySettings(scip, boundstreps, quiet) );
   if( rangedrow )
   {
      /* the gcd rule on, its branches that ADD constraints off (they change the model, not the bounds) */
      SCIP_CALL( SCIPsetBoolParam(scip, "constraints/linear/rangedrowpropagation", TRUE) );
      SCIP_CALL( SCIPsetBoolParam(scip, "constraints/linear/rangedrowartcons", FALSE) );
   }

   t0 = wallclock();
   if( readfile != NULL )
      SCIP_CALL( SCIPreadProb(scip, readfile, NULL) );
   else
   {
      LPB p;
      memset(&p, 0, sizeof(p));
      if( !lpb_read(lpbfile, &p) )
      {
         fprintf(stderr, "ref_driver: cannot read %s\n", lpbfile);
         return SCIP_ERROR;
      }
      SCIP_CALL( buildFromLpb(scip, &p) );
      free(p.rowptr); free(p.colidx); free(p.vals); free(p.lhs); free(p.rhs); free(p.lb); free(p.ub); free(p.vartype);
   }
   tbuild = wallclock() - t0;

   propdata.norigvars = SCIPgetNOrigVars(scip);
   propdata.origvars = (SCIP_VAR**)malloc(sizeof(SCIP_VAR*) * (size_t)(propdata.norigvars + 1));
   memcpy(propdata.origvars, g_lpbvars != NULL ? g_lpbvars : SCIPgetOrigVars(scip), sizeof(SCIP_VAR*) * (size_t)propdata.norigvars);

   /* --repeat K: the model is built once and solved K times (SCIPfreeTransform in between; the statistics of a solve
    * start from zero, misc/resetstat); one REPEAT line per solve but the last, whose results follow as usual */
   for( rep = 0; ; ++rep )
   {
      t0 = wallclock();
      SCIP_CALL( SCIPtransformProb(scip) );
      ttrans = wallclock() - t0;
      SCIP_CALL( SCIPpresolve(scip) );
      tpre = wallclock() - t0 - ttrans;
      SCIP_CALL( SCIPsolve(scip) );
      t1 = wallclock();
      if( !quiet )
         fprintf(stderr, "ref_driver: transform %.3f s, presolve stage %.3f s, solve stage %.3f s\n", ttrans, tpre, t1 - t0 - ttrans - tpre);
      if( rep + 1 >= repeat )
         break;
      printf("REPEAT {\"prop_time_s\": %.9g, \"solve_time_s\": %.9g, \"prop_calls\": %lld, \"domreds\": %lld}\n",
         SCIPconshdlrGetPropTime(SCIPfindConshdlr(scip, "linear")), t1 - t0,
         (long long)SCIPconshdlrGetNPropCalls(SCIPfindConshdlr(scip, "linear")),
         (long long)SCIPconshdlrGetNDomredsFound(SCIPfindConshdlr(scip, "linear")));
      fflush(stdout);
      SCIP_CALL( SCIPfreeTransform(scip) );
   }

   infeasible = (SCIPgetStatus(scip) == SCIP_STATUS_INFEASIBLE);
   if( g_reprop > 0 && g_repropdone )
      infeasible = g_repropcutoff;
   linhdlr = SCIPfindConshdlr(scip, "linear");

   printf("{\"status\": \"%s\", \"ncols\": %d, \"nrows\": %d, \"prop_calls\": %lld, \"domreds\": %lld, "
      "\"prop_time_s\": %.9g, \"solve_time_s\": %.9g, \"build_time_s\": %.9g, \"dumped\": %d}\n",
      infeasible ? "infeasible" : "ok", propdata.norigvars, SCIPconshdlrGetNActiveConss(linhdlr),
      (long long)SCIPconshdlrGetNPropCalls(linhdlr), (long long)SCIPconshdlrGetNDomredsFound(linhdlr),
      SCIPconshdlrGetPropTime(linhdlr), t1 - t0, tbuild, propdata.dumped);

   if( outfile != NULL )
   {
      int64_t ncols = propdata.norigvars;
      int32_t status = infeasible;
      int32_t ncalls = (int32_t)SCIPconshdlrGetNPropCalls(linhdlr);
      int64_t ndomreds = SCIPconshdlrGetNDomredsFound(linhdlr);
      double proptime = SCIPconshdlrGetPropTime(linhdlr);
      double solvetime = t1 - t0;
      f = fopen(outfile, "wb");
      if( f == NULL )
         return SCIP_ERROR;
      fwrite(LPR_MAGIC, 1, 8, f);
      fwrite(&ncols, 8, 1, f);
      fwrite(&status, 4, 1, f);
      fwrite(&ncalls, 4, 1, f);
      fwrite(&ndomreds, 8, 1, f);
      fwrite(&proptime, 8, 1, f);
      fwrite(&solvetime, 8, 1, f);
      for( i = 0; i < propdata.norigvars; ++i )
      {
         SCIP_VAR* tv = SCIPvarGetTransVar(propdata.origvars[i]);
         double b = (tv != NULL) ? SCIPvarGetLbGlobal(tv) : SCIPvarGetLbGlobal(propdata.origvars[i]);
         if( g_reprop > 0 && g_reproplb != NULL )
            b = g_reproplb[i];
         fwrite(&b, 8, 1, f);
      }
      for( i = 0; i < propdata.norigvars; ++i )
      {
         SCIP_VAR* tv = SCIPvarGetTransVar(propdata.origvars[i]);
         double b = (tv != NULL) ? SCIPvarGetUbGlobal(tv) : SCIPvarGetUbGlobal(propdata.origvars[i]);
         if( g_reprop > 0 && g_repropub != NULL )
            b = g_repropub[i];
         fwrite(&b, 8, 1, f);
      }
      fclose(f);
   }

   /* --redundant OUT: one byte per original constraint (= row of the .lpb, in order): 1 if the reference removed it during
    * the root propagation -- propagateCons deletes rows it finds redundant (cons_linear.c:7743-7753; at depth 0
    * SCIPdelConsLocal is a global deletion) */
   if( redfile != NULL )
   {
      SCIP_CONS** oconss = SCIPgetOrigConss(scip);
      int noconss = SCIPgetNOrigConss(scip);
      f = fopen(redfile, "wb");
      if( f == NULL )
         return SCIP_ERROR;
      for( i = 0; i < noconss; ++i )
      {
         SCIP_CONS* tcons = NULL;
         unsigned char gone;
         SCIP_CALL( SCIPgetTransformedCons(scip, oconss[i], &tcons) );
         gone = (unsigned char)((tcons == NULL || SCIPconsIsDeleted(tcons) || !SCIPconsIsActive(tcons)) ? 1 : 0);
         fwrite(&gone, 1, 1, f);
      }
      fclose(f);
   }

   if( g_tiefile != NULL && g_tie != NULL )
   {
      f = fopen(g_tiefile, "wb");
      if( f == NULL )
         return SCIP_ERROR;
      fwrite(g_tie, sizeof(int32_t), (size_t)propdata.norigvars, f);
      fclose(f);
   }
   free(g_tie);
   free(propdata.origvars);
   free(g_probidx2orig);
   free(g_lpbvars);
   free(g_fixvar); free(g_fixval); free(g_reproplb); free(g_repropub);
   SCIP_CALL( SCIPfree(&scip) );
   return SCIP_OKAY;
}

int main(int argc, char** argv)
{
   SCIP_RETCODE rc = run(argc, argv);
   if( rc != SCIP_OKAY )
   {
      fprintf(stderr, "ref_driver: SCIP error %d\n", (int)rc);
      return 1;
   }
   return 0;
}
