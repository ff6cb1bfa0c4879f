/* gpulinear_driver.c -- a minimal SCIP application with prop_gpulinear linked in (SCIP has no dlopen plugin loader:
 * plugins are linked into the application, cf. examples/Eventhdlr/src/cmain.c of the reference).
 *
 *    gpulinear_driver (--lpb X.lpb | --read X.mps) [--cpu] [--boundstreps B] [--out X.lpr] [--solve] [--presolve] [--verbose]
 *
 * Root-node propagation to the fixpoint with presolving / LP / heuristics / all other propagators off.  By default the
 * bound tightening of the linear constraint handler is switched off (constraints/linear/tightenboundsfreq = -1) and
 * prop_gpulinear does the work on the GPU; --cpu leaves everything to the reference (for A/B runs).  --solve lifts the
 * node limit and solves to optimality by propagation + branching (the build has no LP solver).
 * Output: one JSON line; --out writes the global bounds after the root propagation (.lpr, see scip_b200/lpb.py).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <time.h>

#include "scip/scip.h"
#include "scip/scipdefplugins.h"
#include "prop_gpulinear.h"

static double wallclock(void)
{
   struct timespec ts;
   clock_gettime(CLOCK_MONOTONIC, &ts);
   return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static SCIP_VAR** g_vars = NULL;   /* variables in file order */
static int g_nvars = 0;

static SCIP_RETCODE readLpb(SCIP* scip, const char* fn)
{
   FILE* f = fopen(fn, "rb");
   char magic[8];
   int64_t nrows, ncols, nnz, i;
   int64_t* rowptr;
   int32_t* colidx;
   double *vals, *lhs, *rhs, *lb, *ub;
   uint8_t* vartype;
   SCIP_VAR** rowvars;
   char name[64];
   size_t pad;
   int maxlen = 0;

   if( f == NULL || fread(magic, 1, 8, f) != 8 || memcmp(magic, "GPULPB01", 8) != 0 )
   {
      fprintf(stderr, "gpulinear_driver: cannot read %s\n", fn);
      return SCIP_READERROR;
   }
   if( fread(&nrows, 8, 1, f) != 1 || fread(&ncols, 8, 1, f) != 1 || fread(&nnz, 8, 1, f) != 1 )
      return SCIP_READERROR;
   rowptr = (int64_t*)malloc(8 * ((size_t)nrows + 1));
   colidx = (int32_t*)malloc(4 * (size_t)nnz + 8);
   vals = (double*)malloc(8 * (size_t)nnz + 8);
   lhs = (double*)malloc(8 * (size_t)nrows + 8);
   rhs = (double*)malloc(8 * (size_t)nrows + 8);
   lb = (double*)malloc(8 * (size_t)ncols + 8);
   ub = (double*)malloc(8 * (size_t)ncols + 8);
   vartype = (uint8_t*)malloc((size_t)ncols + 8);
   pad = (size_t)((8 - (4 * nnz) % 8) % 8);
   if( fread(rowptr, 8, (size_t)nrows + 1, f) != (size_t)nrows + 1 || fread(colidx, 4, (size_t)nnz, f) != (size_t)nnz
      || fread(magic, 1, pad, f) != pad || fread(vals, 8, (size_t)nnz, f) != (size_t)nnz
      || fread(lhs, 8, (size_t)nrows, f) != (size_t)nrows || fread(rhs, 8, (size_t)nrows, f) != (size_t)nrows
      || fread(lb, 8, (size_t)ncols, f) != (size_t)ncols || fread(ub, 8, (size_t)ncols, f) != (size_t)ncols
      || fread(vartype, 1, (size_t)ncols, f) != (size_t)ncols )
      return SCIP_READERROR;
   fclose(f);

   SCIP_CALL( SCIPcreateProbBasic(scip, "lpb") );
   g_vars = (SCIP_VAR**)malloc(sizeof(SCIP_VAR*) * ((size_t)ncols + 1));
   g_nvars = (int)ncols;
   for( i = 0; i < ncols; ++i )
   {
      SCIP_VARTYPE vt = SCIP_VARTYPE_CONTINUOUS;
      if( vartype[i] )
         vt = (lb[i] == 0.0 && ub[i] == 1.0) ? SCIP_VARTYPE_BINARY : SCIP_VARTYPE_INTEGER;
      snprintf(name, sizeof(name), "x%lld", (long long)i);
      SCIP_CALL( SCIPcreateVarBasic(scip, &g_vars[i], name, lb[i], ub[i], 0.0, vt) );
      SCIP_CALL( SCIPaddVar(scip, g_vars[i]) );
   }
   for( i = 0; i < nrows; ++i )
      if( rowptr[i + 1] - rowptr[i] > maxlen )
         maxlen = (int)(rowptr[i + 1] - rowptr[i]);
   rowvars = (SCIP_VAR**)malloc(sizeof(SCIP_VAR*) * ((size_t)maxlen + 1));
   for( i = 0; i < nrows; ++i )
   {
      SCIP_CONS* cons;
      int len = (int)(rowptr[i + 1] - rowptr[i]);
      int v;
      for( v = 0; v < len; ++v )
         rowvars[v] = g_vars[colidx[rowptr[i] + v]];
      snprintf(name, sizeof(name), "c%lld", (long long)i);
      SCIP_CALL( SCIPcreateConsBasicLinear(scip, &cons, name, len, rowvars, vals + rowptr[i], lhs[i], rhs[i]) );
      SCIP_CALL( SCIPaddCons(scip, cons) );
      SCIP_CALL( SCIPreleaseCons(scip, &cons) );
   }
   free(rowvars); free(rowptr); free(colidx); free(vals); free(lhs); free(rhs); free(lb); free(ub); free(vartype);
   return SCIP_OKAY;
}

/* --probe-batch N: a checker propagator (delayed: it runs once the root node is propagated) probes up to N unfixed
 * integral variables the way SCIPapplyProbingVar does (prop_probing.c:1254-1279: SCIPstartProbing, SCIPchgVarLb/UbProbing,
 * SCIPpropagateProbing, SCIPendProbing), then hands the same probes to SCIPprobeBatchGpulinear and compares the verdicts */
static int g_nprobecheck = 0;
static int g_probedone = 0;
static int g_inprobecheck = 0;
static int g_usecpu = 0;

static SCIP_DECL_PROPEXEC(propExecProbecheck)
{
   SCIP_VAR** vars = SCIPgetVars(scip);
   SCIP_VAR** pv;
   SCIP_Real* plb;
   SCIP_Real* pub;
   SCIP_Bool* cutscip;
   SCIP_Bool* cutbatch;
   SCIP_Longint* ndom;
   SCIP_Longint* nchg;
   SCIP_Real* probelb;        /* local bounds after SCIP's own probing cycle, n x nvars */
   SCIP_Real* probeub;
   SCIP_Real* nodelb;
   SCIP_Real* nodeub;
   int* chgbeg;
   SCIP_VAR** chgvars;
   SCIP_BOUNDTYPE* chgtypes;
   SCIP_Real* chgbounds;
   int nchgs = 0;
   int maxchgs;
   int nboundmismatch = 0;
   int v;
   SCIP_Bool nodecutoff = FALSE;
   int nvars = SCIPgetNVars(scip);
   int n = 0;
   int nmismatch = 0;
   int ncutscip = 0;
   int ncutbatch = 0;
   int nactive = 0;
   double t0, t1, t2;
   int i;

   (void)prop;
   (void)proptiming;
   *result = SCIP_DIDNOTRUN;
   if( g_probedone || g_inprobecheck || SCIPgetDepth(scip) != 0 || SCIPinProbing(scip) )
      return SCIP_OKAY;
   g_probedone = 1;
   g_inprobecheck = 1;
   SCIP_CALL( SCIPallocBufferArray(scip, &pv, g_nprobecheck) );
   SCIP_CALL( SCIPallocBufferArray(scip, &plb, g_nprobecheck) );
   SCIP_CALL( SCIPallocBufferArray(scip, &pub, g_nprobecheck) );
   SCIP_CALL( SCIPallocBufferArray(scip, &cutscip, g_nprobecheck) );
   SCIP_CALL( SCIPallocBufferArray(scip, &cutbatch, g_nprobecheck) );
   SCIP_CALL( SCIPallocBufferArray(scip, &ndom, g_nprobecheck) );
   SCIP_CALL( SCIPallocBufferArray(scip, &nchg, g_nprobecheck) );
   for( i = 0; i < nvars && n < g_nprobecheck; ++i )
   {
      const SCIP_Real lb = SCIPvarGetLbLocal(vars[i]);
      const SCIP_Real ub = SCIPvarGetUbLocal(vars[i]);
      if( !SCIPvarIsIntegral(vars[i]) || ub - lb < 0.5 || SCIPisInfinity(scip, -lb) || SCIPisInfinity(scip, ub) )
         continue;
      pv[n] = vars[i];
      if( n % 2 == 0 )
      {
         /* upper part of the domain */
         plb[n] = SCIPfeasFloor(scip, 0.5 * (lb + ub)) + 1.0;
         pub[n] = ub;
      }
      else
      {
         plb[n] = lb;
         pub[n] = SCIPfeasFloor(scip, 0.5 * (lb + ub));
      }
      ++n;
   }
   maxchgs = 64 * (n + 1) + 4 * nvars;
   SCIP_CALL( SCIPallocBufferArray(scip, &probelb, (n + 1) * nvars) );
   SCIP_CALL( SCIPallocBufferArray(scip, &probeub, (n + 1) * nvars) );
   SCIP_CALL( SCIPallocBufferArray(scip, &nodelb, nvars) );
   SCIP_CALL( SCIPallocBufferArray(scip, &nodeub, nvars) );
   SCIP_CALL( SCIPallocBufferArray(scip, &chgbeg, n + 1) );
   SCIP_CALL( SCIPallocBufferArray(scip, &chgvars, maxchgs) );
   SCIP_CALL( SCIPallocBufferArray(scip, &chgtypes, maxchgs) );
   SCIP_CALL( SCIPallocBufferArray(scip, &chgbounds, maxchgs) );
   t0 = wallclock();
   for( i = 0; i < n; ++i )
   {
      SCIP_CALL( SCIPstartProbing(scip) );
      if( plb[i] > SCIPvarGetLbLocal(pv[i]) )
         SCIP_CALL( SCIPchgVarLbProbing(scip, pv[i], plb[i]) );
      if( pub[i] < SCIPvarGetUbLocal(pv[i]) )
         SCIP_CALL( SCIPchgVarUbProbing(scip, pv[i], pub[i]) );
      SCIP_CALL( SCIPpropagateProbing(scip, -1, &cutscip[i], &ndom[i]) );
      for( v = 0; v < nvars; ++v )
      {
         probelb[i * nvars + v] = SCIPvarGetLbLocal(vars[v]);
         probeub[i * nvars + v] = SCIPvarGetUbLocal(vars[v]);
      }
      SCIP_CALL( SCIPendProbing(scip) );
      ncutscip += cutscip[i] ? 1 : 0;
   }
   t1 = wallclock();
   t2 = t1;
   if( !g_usecpu )
   {
      SCIP_CALL( SCIPprobeBatchGpulinear(scip, n, pv, plb, pub, &nodecutoff, cutbatch, NULL, nchg) );
      t2 = wallclock();
      for( i = 0; i < n; ++i )
      {
         ncutbatch += cutbatch[i] ? 1 : 0;
         nactive += nchg[i] > 0 ? 1 : 0;
         if( cutbatch[i] != cutscip[i] )
            ++nmismatch;
      }
      /* the implied bounds of every probe (SCIPapplyProbingVar's proplbs / propubs): the node's bounds overwritten by the
       * probe's bound changes must be the bounds SCIP's own probing cycle ended with */
      SCIP_CALL( SCIPprobeBatchBoundsGpulinear(scip, n, pv, plb, pub, &nodecutoff, cutbatch, chgbeg, chgvars, chgtypes,
            chgbounds, maxchgs, &nchgs) );
      for( v = 0; v < nvars; ++v )
      {
         nodelb[v] = SCIPvarGetLbLocal(vars[v]);
         nodeub[v] = SCIPvarGetUbLocal(vars[v]);
      }
      for( i = 0; i < n && nchgs <= maxchgs; ++i )
      {
         int e;
         if( cutbatch[i] || cutscip[i] )
            continue;
         for( v = 0; v < nvars; ++v )
         {
            probelb[n * nvars + v] = nodelb[v];
            probeub[n * nvars + v] = nodeub[v];
         }
         v = SCIPvarGetProbindex(pv[i]);
         probelb[n * nvars + v] = MAX(nodelb[v], plb[i]);
         probeub[n * nvars + v] = MIN(nodeub[v], pub[i]);
         for( e = chgbeg[i]; e < chgbeg[i + 1]; ++e )
         {
            v = SCIPvarGetProbindex(chgvars[e]);
            if( chgtypes[e] == SCIP_BOUNDTYPE_UPPER )
               probeub[n * nvars + v] = chgbounds[e];
            else
               probelb[n * nvars + v] = chgbounds[e];
         }
         for( v = 0; v < nvars; ++v )
         {
            if( !SCIPisFeasEQ(scip, probelb[n * nvars + v], probelb[i * nvars + v])
               || !SCIPisFeasEQ(scip, probeub[n * nvars + v], probeub[i * nvars + v]) )
               ++nboundmismatch;
         }
      }
   }
   printf("PROBEBATCH {\"probes\": %d, \"cutoffs_scip\": %d, \"cutoffs_batch\": %d, \"mismatches\": %d, \"node_cutoff\": %d, "
      "\"probes_with_changes\": %d, \"implied_bound_changes\": %d, \"implied_bound_mismatches\": %d, "
      "\"scip_probing_s\": %.9g, \"batch_s\": %.9g}\n", n, ncutscip, ncutbatch, nmismatch,
      (int)nodecutoff, nactive, nchgs, nboundmismatch, t1 - t0, t2 - t1);
   SCIPfreeBufferArray(scip, &chgbounds);
   SCIPfreeBufferArray(scip, &chgtypes);
   SCIPfreeBufferArray(scip, &chgvars);
   SCIPfreeBufferArray(scip, &chgbeg);
   SCIPfreeBufferArray(scip, &nodeub);
   SCIPfreeBufferArray(scip, &nodelb);
   SCIPfreeBufferArray(scip, &probeub);
   SCIPfreeBufferArray(scip, &probelb);
   SCIPfreeBufferArray(scip, &nchg);
   SCIPfreeBufferArray(scip, &ndom);
   SCIPfreeBufferArray(scip, &cutbatch);
   SCIPfreeBufferArray(scip, &cutscip);
   SCIPfreeBufferArray(scip, &pub);
   SCIPfreeBufferArray(scip, &plb);
   SCIPfreeBufferArray(scip, &pv);
   g_inprobecheck = 0;
   *result = SCIP_DIDNOTFIND;
   return SCIP_OKAY;
}

static SCIP_RETCODE run(int argc, char** argv)
{
   static const char* offprops[] = { "dualfix", "genvbounds", "nlobbt", "obbt", "probing", "pseudoobj", "redcost",
      "rootredcost", "vbounds", "symmetry", NULL };
   SCIP* scip = NULL;
   SCIP_CONSHDLR* linhdlr;
   SCIP_PROP* gpuprop;
   const char* readfile = NULL;
   const char* lpbfile = NULL;
   const char* outfile = NULL;
   double boundstreps = -1.0;
   int usecpu = 0;
   int solve = 0;
   int quiet = 1;
   int presolve = 0;
   int activeonly = 0;
   int delredundant = 0;
   int ndevices = 1;
   int rangedrow = 0;
   int rowsof[5] = {-1, -1, -1, -1, -1};
   char pname[128];
   double t0, t1;
   int infeasible;
   int i;

   for( i = 1; i < argc; ++i )
   {
      if( strcmp(argv[i], "--read") == 0 && i + 1 < argc ) readfile = argv[++i];
      else if( strcmp(argv[i], "--lpb") == 0 && i + 1 < argc ) lpbfile = argv[++i];
      else if( strcmp(argv[i], "--out") == 0 && i + 1 < argc ) outfile = argv[++i];
      else if( strcmp(argv[i], "--boundstreps") == 0 && i + 1 < argc ) boundstreps = atof(argv[++i]);
      else if( strcmp(argv[i], "--cpu") == 0 ) usecpu = 1;
      else if( strcmp(argv[i], "--solve") == 0 ) solve = 1;
      else if( strcmp(argv[i], "--verbose") == 0 ) quiet = 0;
      else if( strcmp(argv[i], "--presolve") == 0 ) presolve = 1;
      else if( strcmp(argv[i], "--active-rows-only") == 0 ) activeonly = 1;
      else if( strcmp(argv[i], "--del-redundant") == 0 ) delredundant = 1;
      else if( strcmp(argv[i], "--ndevices") == 0 && i + 1 < argc ) ndevices = atoi(argv[++i]);
      else if( strcmp(argv[i], "--rangedrow") == 0 ) rangedrow = 1;
      else if( strcmp(argv[i], "--probe-batch") == 0 && i + 1 < argc ) g_nprobecheck = atoi(argv[++i]);
      else
      {
         fprintf(stderr, "usage: gpulinear_driver (--lpb F | --read F) [--cpu] [--boundstreps B] [--out F.lpr] [--solve] [--presolve] [--verbose]\n");
         return SCIP_ERROR;
      }
   }
   if( (readfile == NULL) == (lpbfile == NULL) )
      return SCIP_ERROR;

   SCIP_CALL( SCIPcreate(&scip) );
   SCIP_CALL( SCIPincludeDefaultPlugins(scip) );
   if( !usecpu )
      SCIP_CALL( SCIPincludePropGpulinear(scip) );
   g_usecpu = usecpu;
   if( g_nprobecheck > 0 )
   {
      SCIP_PROP* checkprop = NULL;
      SCIP_CALL( SCIPincludePropBasic(scip, &checkprop, "probecheck", "compares SCIP's probing cycle with the batched GPU entry",
            -1000, 1, TRUE, SCIP_PROPTIMING_BEFORELP, propExecProbecheck, NULL) );
   }

   if( !presolve )
   {
      SCIP_CALL( SCIPsetIntParam(scip, "presolving/maxrounds", 0) );
      SCIP_CALL( SCIPsetIntParam(scip, "presolving/maxrestarts", 0) );
   }
   SCIP_CALL( SCIPsetIntParam(scip, "propagating/maxrounds", -1) );
   SCIP_CALL( SCIPsetIntParam(scip, "propagating/maxroundsroot", -1) );
   SCIP_CALL( SCIPsetIntParam(scip, "lp/solvefreq", -1) );
   if( !solve )
      SCIP_CALL( SCIPsetLongintParam(scip, "limits/nodes", 1LL) );
   SCIP_CALL( SCIPsetBoolParam(scip, "conflict/enable", solve ? TRUE : FALSE) );
   /* --rangedrow: the gcd rule on -- in cons_linear with --cpu, else on the device; never its artificial constraints */
   SCIP_CALL( SCIPsetBoolParam(scip, "constraints/linear/rangedrowpropagation", (usecpu && rangedrow) ? TRUE : FALSE) );
   SCIP_CALL( SCIPsetBoolParam(scip, "constraints/linear/rangedrowartcons", FALSE) );
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
   if( !usecpu )
   {
      SCIP_CALL( SCIPsetIntParam(scip, "constraints/linear/tightenboundsfreq", -1) );   /* the replaced path is off */
      if( activeonly )
         SCIP_CALL( SCIPsetBoolParam(scip, "propagating/gpulinear/stablecopy", FALSE) );
      if( delredundant )
         SCIP_CALL( SCIPsetBoolParam(scip, "propagating/gpulinear/delredundant", TRUE) );
      if( ndevices > 1 )
         SCIP_CALL( SCIPsetIntParam(scip, "propagating/gpulinear/ndevices", ndevices) );
      if( rangedrow )
         SCIP_CALL( SCIPsetBoolParam(scip, "propagating/gpulinear/rangedrow", TRUE) );
   }

   if( readfile != NULL )
      SCIP_CALL( SCIPreadProb(scip, readfile, NULL) );
   else
      SCIP_CALL( readLpb(scip, lpbfile) );
   if( g_vars == NULL )
   {
      g_nvars = SCIPgetNOrigVars(scip);
      g_vars = (SCIP_VAR**)malloc(sizeof(SCIP_VAR*) * ((size_t)g_nvars + 1));
      memcpy(g_vars, SCIPgetOrigVars(scip), sizeof(SCIP_VAR*) * (size_t)g_nvars);
   }

   t0 = wallclock();
   SCIP_CALL( SCIPsolve(scip) );
   t1 = wallclock();

   infeasible = (SCIPgetStatus(scip) == SCIP_STATUS_INFEASIBLE);
   if( !usecpu )
   {
      for( i = 0; i < 5; ++i )
         rowsof[i] = SCIPgetNRowsGpulinear(scip, i);
   }
   linhdlr = SCIPfindConshdlr(scip, "linear");
   gpuprop = SCIPfindProp(scip, "gpulinear");
   printf("{\"mode\": \"%s\", \"status\": \"%s\", \"scip_status\": %d, \"ncols\": %d, \"nodes\": %lld, \"linear_prop_calls\": %lld, "
      "\"linear_domreds\": %lld, \"linear_prop_time_s\": %.9g, \"gpu_prop_calls\": %lld, \"gpu_domreds\": %lld, "
      "\"gpu_prop_time_s\": %.9g, \"solve_time_s\": %.9g, \"primal\": %.15g, \"gpu_rows\": [%d, %d, %d, %d, %d], \"gpu_builds\": %lld}\n",
      usecpu ? "cpu" : "gpu", infeasible ? "infeasible" : "ok", (int)SCIPgetStatus(scip), g_nvars, (long long)SCIPgetNNodes(scip),
      (long long)SCIPconshdlrGetNPropCalls(linhdlr), (long long)SCIPconshdlrGetNDomredsFound(linhdlr),
      SCIPconshdlrGetPropTime(linhdlr), gpuprop != NULL ? (long long)SCIPpropGetNCalls(gpuprop) : 0LL,
      gpuprop != NULL ? (long long)SCIPpropGetNDomredsFound(gpuprop) : 0LL, gpuprop != NULL ? SCIPpropGetTime(gpuprop) : 0.0,
      t1 - t0, SCIPgetPrimalbound(scip), rowsof[0], rowsof[1], rowsof[2], rowsof[3], rowsof[4],
      usecpu ? 0LL : (long long)SCIPgetNBuildsGpulinear(scip));

   if( outfile != NULL )
   {
      FILE* f = fopen(outfile, "wb");
      int64_t ncols = g_nvars;
      int32_t status = infeasible;
      int32_t ncalls = (int32_t)(gpuprop != NULL ? SCIPpropGetNCalls(gpuprop) : SCIPconshdlrGetNPropCalls(linhdlr));
      int64_t ndomreds = gpuprop != NULL ? SCIPpropGetNDomredsFound(gpuprop) : SCIPconshdlrGetNDomredsFound(linhdlr);
      double proptime = gpuprop != NULL ? SCIPpropGetTime(gpuprop) : SCIPconshdlrGetPropTime(linhdlr);
      double solvetime = t1 - t0;
      if( f == NULL )
         return SCIP_ERROR;
      fwrite("GPULPR01", 1, 8, f);
      fwrite(&ncols, 8, 1, f);
      fwrite(&status, 4, 1, f);
      fwrite(&ncalls, 4, 1, f);
      fwrite(&ndomreds, 8, 1, f);
      fwrite(&proptime, 8, 1, f);
      fwrite(&solvetime, 8, 1, f);
      for( i = 0; i < g_nvars; ++i )
      {
         SCIP_VAR* tv = SCIPvarGetTransVar(g_vars[i]);
         double b = SCIPvarGetLbGlobal(tv != NULL ? tv : g_vars[i]);
         fwrite(&b, 8, 1, f);
      }
      for( i = 0; i < g_nvars; ++i )
      {
         SCIP_VAR* tv = SCIPvarGetTransVar(g_vars[i]);
         double b = SCIPvarGetUbGlobal(tv != NULL ? tv : g_vars[i]);
         fwrite(&b, 8, 1, f);
      }
      fclose(f);
   }
   if( lpbfile != NULL )
   {
      for( i = 0; i < g_nvars; ++i )
         SCIP_CALL( SCIPreleaseVar(scip, &g_vars[i]) );
   }
   free(g_vars);
   SCIP_CALL( SCIPfree(&scip) );
   return SCIP_OKAY;
}

int main(int argc, char** argv)
{
   SCIP_RETCODE rc = run(argc, argv);
   if( rc != SCIP_OKAY )
   {
      fprintf(stderr, "gpulinear_driver: SCIP error %d\n", (int)rc);
      return 1;
   }
   return 0;
}
