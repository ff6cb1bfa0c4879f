/* prop_gpulinear.c -- B200 linear bound propagation as an ordinary SCIP propagator plugin.
 *
 * Drop-in for the path  propagationRound (solve.c:437) -> consPropLinear (cons_linear.c:16126) -> propagateCons (:7621)
 * -> tightenBounds (:6980).  As a propagator with non-negative priority it is called by SCIPpropExec (prop.c:646) before
 * the constraint handlers (solve.c:524).  What it does per call (SCIP_DECL_PROPEXEC, type_prop.h:217):
 *
 *   1. (re)build the device copy of the linear rows when the set of constraints changed: rows through
 *      SCIPgetVarsLinear / SCIPgetValsLinear / SCIPgetLhsLinear / SCIPgetRhsLinear (cons_linear.h:263-307) and -- with
 *      propagating/gpulinear/allrows, the way SCIP's own matrix view does it (matrix.c:541-696) -- the linear rows
 *      behind knapsack, setppc, logicor and varbound constraints, into which presolving upgrades most linear
 *      constraints; every row is rewritten to active variables (SCIPgetProbvarLinearSum, scip_var.h:962: negated and
 *      aggregated variables, the constant goes to the sides), columns = SCIPvarGetProbindex (pub_var.h:592); tolerances from SCIPinfinity/SCIPepsilon/SCIPsumepsilon/SCIPfeastol/
 *      SCIPgetHugeValue and numerics/boundstreps, constraints/linear/maxeasyactivitydelta;
 *   2. bring the device bounds up to date: the first call after a (re)build sends all local bounds
 *      (SCIPvarGetLbLocal/UbLocal, gpulin_set_bounds); later calls send only the bounds that changed since the last
 *      call -- an event handler catches SCIP_EVENTTYPE_BOUNDCHANGED of every variable (SCIPcatchVarEvent,
 *      scip_event.h:254), exactly as cons_linear does for its rows (cons_linear.c:717-754, eventExecLinear :17162) --
 *      through gpulin_update_bounds, which marks only the rows of those columns (cost proportional to the changes:
 *      branching, backtracking, other propagators);
 *   3. gpulin_propagate: all rounds to the fixpoint on the GPU (no host round trip per round);
 *   4. replay the round-ordered change log with SCIPinferVarLbProp / SCIPinferVarUbProp (scip_var.c:7589/7705,
 *      force = TRUE: the relative threshold numerics/boundstreps was already applied on the device per inference);
 *      the replay order makes every change explainable by bounds that SCIP already knows (SCIP_DECL_PROPRESPROP);
 *   5. *result = SCIP_CUTOFF / SCIP_REDUCEDDOM / SCIP_DIDNOTFIND (type_result.h:42-51).
 *
 * There is no CPU fallback: a CUDA failure is returned as SCIP_ERROR.  The propagator is not copied to sub-SCIPs
 * (PROPCOPY = NULL): copies do not get a GPU.
 */
#include <assert.h>
#include <stdlib.h>
#include <string.h>

#include "prop_gpulinear.h"
#include "gpulin.h"

#include "scip/cons_linear.h"
#include "scip/cons_knapsack.h"
#include "scip/cons_logicor.h"
#include "scip/cons_setppc.h"
#include "scip/cons_varbound.h"
#include "scip/pub_cons.h"
#include "scip/pub_event.h"
#include "scip/scip_event.h"
#include "scip/pub_message.h"
#include "scip/pub_prop.h"
#include "scip/pub_var.h"
#include "scip/scip_conflict.h"
#include "scip/scip_cons.h"
#include "scip/scip_general.h"
#include "scip/scip_mem.h"
#include "scip/scip_message.h"
#include "scip/scip_numerics.h"
#include "scip/scip_param.h"
#include "scip/scip_prob.h"
#include "scip/scip_prop.h"
#include "scip/scip_probing.h"
#include "scip/scip_tree.h"
#include "scip/scip_var.h"

#define PROP_NAME              "gpulinear"
#define PROP_DESC              "activity-based bound propagation of all linear constraints on a B200 GPU (libgpulin)"
#define PROP_PRIORITY          10000000   /* >= 0: before the constraint handlers (solve.c:524) */
#define PROP_FREQ              1
#define PROP_DELAY             FALSE
#define PROP_TIMING            SCIP_PROPTIMING_BEFORELP

#define DEFAULT_MAXROUNDS      -1         /* rounds per call on the device (-1: to the fixpoint) */
#define DEFAULT_DEVICE         0
#define DEFAULT_LOGCAPFAC      8          /* change log capacity = factor * number of variables */
#define DEFAULT_DELREDUNDANT   FALSE      /* delete rows locally that the device finds redundant (cons_linear does so itself for its rows) */
#define DEFAULT_DETERMINISTIC  TRUE       /* replay the changes of a round in a fixed order */
#define DEFAULT_STABLECOPY     TRUE       /* device copy of all existing global rows, not only of those active at the node of the build */
#define DEFAULT_ALLROWS        TRUE       /* read the rows of knapsack / setppc / logicor / varbound constraints as well */
#define DEFAULT_INCREMENTAL    TRUE       /* send only changed bounds (event driven) instead of all bounds per call */
#define EVENTHDLR_NAME         "gpulinear"
#define EVENTHDLR_DESC         "collects the variables whose bounds changed since the last call of prop_gpulinear"

struct SCIP_PropData
{
   gpulin_t*             gpu;                /**< device copy of the linear rows, or NULL */
   gpulin_t*             peergpu[7];         /**< ndevices > 1: the copies on the other devices (gpulin_group_connect) */
   int                   npeergpus;          /**< copies in peergpu[] */
   int                   ndevices;           /**< parameter: devices that share the dense rounds (device, device + 1, ...) */
   SCIP_VAR**            vars;               /**< column -> variable (by probindex at build time) */
   SCIP_CONS**           rowcons;            /**< row -> linear constraint */
   SCIP_Real*            lb;                 /**< host bound buffers */
   SCIP_Real*            ub;
   SCIP_Real*            reflb;              /**< reference bounds on the device (global bounds at build time): a full bound */
   SCIP_Real*            refub;              /**<   upload sends 2 bits per column against them (gpulin_set_bounds_packed) */
   uint32_t*             codes;              /**< the 2-bit codes, 16 columns per word */
   gpulin_change*        changes;            /**< change log buffer */
   uint32_t*             clog;               /**< long logs come back compact (gpulin_get_changes_compact): one word per entry ... */
   uint32_t*             cside;              /**<   ... and three per entry whose bound is not 0 or 1 */
   int64_t*              rowptr;             /**< host CSR (kept for PROPRESPROP) */
   int32_t*              colidx;
   SCIP_Real*            vals;
   SCIP_Real*            lhs;
   SCIP_Real*            rhs;
   int64_t*              colptr;             /**< column -> rows (for PROPRESPROP) */
   int32_t*              colrows;
   int32_t*              colpos;             /**< position of the entry in its row */
   int                   ncols;
   int                   nrows;
   int64_t               nnz;
   int64_t               logcap;
   int                   nlinconss;          /**< active constraints of all row sources when the device copy was built */
   SCIP_CONS*            sourcetail[5];      /**< last constraint in the array of every row source at build time (new constraints
                                              *   are appended, a deletion moves the last one: part of the staleness check) */
   int                   nrowsof[5];         /**< rows per source: linear, knapsack, setppc, logicor, varbound */
   SCIP_Bool             allrows;            /**< parameter: also read knapsack / setppc / logicor / varbound rows */
   SCIP_Bool             delredundant;       /**< parameter: SCIPdelConsLocal for rows the device proves redundant */
   SCIP_Bool             rangedrow;          /**< parameter: ranged-row (gcd) propagation on the device (constraints/linear/rangedrowpropagation) */
   SCIP_Longint          ndelconss;          /**< constraints deleted locally on the device's verdict */
   int32_t*              redrows;            /**< buffer for gpulin_get_redundant_rows */
   SCIP_Bool             deterministic;      /**< parameter: sort the change log inside every round before the replay */
   SCIP_Bool             stablecopy;         /**< parameter: the device copy holds every existing global row (see countSourceConss) */
   SCIP_Longint          nbuilds;            /**< device copies built so far */
   int                   nskipped;           /**< rows not sent to the device (modifiable, local, non-active variables) */
   SCIP_EVENTHDLR*       eventhdlr;          /**< bound change event handler */
   int*                  filterpos;          /**< per column: position of the caught event, or -1 */
   int*                  touched;            /**< columns whose bounds changed since the last call */
   SCIP_Bool*            istouched;          /**< per column: already in touched[] */
   int32_t*              updidx;             /**< staging of gpulin_update_bounds */
   int                   ntouched;
   SCIP_Bool             fullsync;           /**< the next call must send all bounds */
   SCIP_Bool             inreplay;           /**< bound changes are our own: the device already has them */
   SCIP_Bool             incremental;        /**< parameter */
   SCIP_Longint          nfullsyncs;
   SCIP_Longint          nupdates;           /**< bounds sent through gpulin_update_bounds */
   int                   maxrounds;          /**< parameter */
   int                   device;             /**< parameter */
   SCIP_Longint          ncalls;
   SCIP_Longint          nrounds;
   SCIP_Longint          nchanges;
   SCIP_Real             devicems;
};

/** frees the device copy and all host arrays */
static
void freeDeviceCopy(
   SCIP*                 scip,
   SCIP_PROPDATA*        propdata
   )
{
   while( propdata->npeergpus > 0 )
   {
      gpulin_destroy(propdata->peergpu[--propdata->npeergpus]);
      propdata->peergpu[propdata->npeergpus] = NULL;
   }
   if( propdata->gpu != NULL )
   {
      gpulin_destroy(propdata->gpu);
      propdata->gpu = NULL;
   }
   if( propdata->filterpos != NULL )
   {
      int j;
      for( j = 0; j < propdata->ncols; ++j )
      {
         if( propdata->filterpos[j] >= 0 )
         {
            SCIP_RETCODE rc = SCIPdropVarEvent(scip, propdata->vars[j], SCIP_EVENTTYPE_BOUNDCHANGED, propdata->eventhdlr,
               (SCIP_EVENTDATA*)(size_t)(j + 1), propdata->filterpos[j]);
            if( rc != SCIP_OKAY )
               SCIPwarningMessage(scip, "prop_gpulinear: could not drop the bound change event of column %d\n", j);
         }
      }
   }
   SCIPfreeBlockMemoryArrayNull(scip, &propdata->filterpos, propdata->ncols);
   SCIPfreeBlockMemoryArrayNull(scip, &propdata->touched, propdata->ncols);
   SCIPfreeBlockMemoryArrayNull(scip, &propdata->istouched, propdata->ncols);
   SCIPfreeBlockMemoryArrayNull(scip, &propdata->updidx, propdata->ncols);
   propdata->ntouched = 0;
   SCIPfreeBlockMemoryArrayNull(scip, &propdata->vars, propdata->ncols);
   SCIPfreeBlockMemoryArrayNull(scip, &propdata->lb, propdata->ncols);
   SCIPfreeBlockMemoryArrayNull(scip, &propdata->ub, propdata->ncols);
   SCIPfreeBlockMemoryArrayNull(scip, &propdata->reflb, propdata->ncols);
   SCIPfreeBlockMemoryArrayNull(scip, &propdata->refub, propdata->ncols);
   SCIPfreeBlockMemoryArrayNull(scip, &propdata->codes, (propdata->ncols + 15) / 16);
   SCIPfreeBlockMemoryArrayNull(scip, &propdata->colptr, propdata->ncols + 1);
   SCIPfreeBlockMemoryArrayNull(scip, &propdata->redrows, propdata->nrows);
   if( propdata->rowcons != NULL )
   {
      /* the rows' constraints were captured at build time (they cannot be freed under the device copy) */
      int r;
      for( r = 0; r < propdata->nrows; ++r )
      {
         if( propdata->rowcons[r] != NULL && SCIPreleaseCons(scip, &propdata->rowcons[r]) != SCIP_OKAY )
            SCIPwarningMessage(scip, "prop_gpulinear: could not release the constraint of row %d\n", r);
      }
   }
   SCIPfreeBlockMemoryArrayNull(scip, &propdata->rowcons, propdata->nrows);
   SCIPfreeBlockMemoryArrayNull(scip, &propdata->lhs, propdata->nrows);
   SCIPfreeBlockMemoryArrayNull(scip, &propdata->rhs, propdata->nrows);
   SCIPfreeBlockMemoryArrayNull(scip, &propdata->rowptr, propdata->nrows + 1);
   SCIPfreeBlockMemoryArrayNull(scip, &propdata->colidx, propdata->nnz);
   SCIPfreeBlockMemoryArrayNull(scip, &propdata->vals, propdata->nnz);
   SCIPfreeBlockMemoryArrayNull(scip, &propdata->colrows, propdata->nnz);
   SCIPfreeBlockMemoryArrayNull(scip, &propdata->colpos, propdata->nnz);
   SCIPfreeBlockMemoryArrayNull(scip, &propdata->changes, propdata->logcap);
   SCIPfreeBlockMemoryArrayNull(scip, &propdata->clog, propdata->logcap);
   SCIPfreeBlockMemoryArrayNull(scip, &propdata->cside, 3 * propdata->logcap);
   propdata->ncols = 0;
   propdata->nrows = 0;
   propdata->nnz = 0;
   propdata->logcap = 0;
   propdata->nlinconss = -1;
}

#define NROWSOURCES 5
static const char* const rowsourcenames[NROWSOURCES] = {"linear", "knapsack", "setppc", "logicor", "varbound"};

/** number of constraints of all row sources (the cheap staleness check of the device copy).
 *
 *  stable = FALSE: the ACTIVE constraints.  In a tree search that number moves from node to node: the handlers delete
 *  constraints locally that have become redundant in a subtree (cons_linear does so in propagateCons, :7743-7753, also when
 *  its bound tightening is off) and backtracking brings them back -- the device copy would be rebuilt at most nodes.
 *  stable = TRUE: the EXISTING constraints (active or not, SCIPconshdlrGetNConss).  The device copy then holds every
 *  existing row that is not deleted, not local and not modifiable.  Such a row is globally valid whether or not it is
 *  active at the current node, so propagating it is always correct; a row that is only switched off because it is
 *  redundant in this subtree cannot tighten anything anyway.  The number changes only when constraints are created or
 *  freed (conflict constraints, restarts). */
static
int countSourceConss(
   SCIP*                 scip,
   SCIP_Bool             allrows,
   SCIP_Bool             stable,
   SCIP_CONS**           tails               /**< NROWSOURCES entries: last constraint of every source's array, or NULL */
   )
{
   int n = 0;
   int s;
   for( s = 0; s < NROWSOURCES; ++s )
   {
      SCIP_CONSHDLR* conshdlr = (allrows || s == 0) ? SCIPfindConshdlr(scip, rowsourcenames[s]) : NULL;
      int ns = 0;
      if( conshdlr != NULL )
         ns = stable ? SCIPconshdlrGetNConss(conshdlr) : SCIPconshdlrGetNActiveConss(conshdlr);
      n += ns;
      if( tails != NULL )
         tails[s] = ns > 0 ? SCIPconshdlrGetConss(conshdlr)[ns - 1] : NULL;
   }
   return n;
}

/** the device copy is stale if the number of source constraints or of variables changed, or -- one constraint created
 *  and another one freed leave the number alone -- if the last constraint of a source's array is a different one (new
 *  constraints are appended, cons.c: conshdlrAddCons; a deletion moves the last constraint into the gap) */
static
SCIP_Bool deviceCopyIsStale(
   SCIP*                 scip,
   SCIP_PROPDATA*        propdata,
   int*                  nsourceconss
   )
{
   SCIP_CONS* tails[NROWSOURCES];
   int s;

   *nsourceconss = countSourceConss(scip, propdata->allrows, propdata->stablecopy, tails);
   if( propdata->gpu == NULL || propdata->nlinconss != *nsourceconss || propdata->ncols != SCIPgetNVars(scip) )
      return TRUE;
   for( s = 0; s < NROWSOURCES; ++s )
   {
      if( tails[s] != propdata->sourcetail[s] )
         return TRUE;
   }
   return FALSE;
}

/** the row  lhs <= sum vals[i] vars[i] <= rhs  of a constraint of source `src`, rewritten to active variables; the
 *  buffers grow as needed.  *usable = FALSE for rows the device must not propagate (cf. tightenBounds :7010: modifiable
 *  rows are skipped; local rows are only valid in a subtree) */
static
SCIP_RETCODE getActiveRow(
   SCIP*                 scip,
   SCIP_CONS*            cons,
   int                   src,
   SCIP_VAR***           vars,
   SCIP_Real**           vals,
   int*                  size,
   int*                  nvars,
   SCIP_Real*            lhs,
   SCIP_Real*            rhs,
   SCIP_Bool*            usable
   )
{
   SCIP_VAR** cvars = NULL;
   SCIP_Real constant = 0.0;
   int n = 0;
   int required = 0;
   int v;

   *usable = FALSE;
   *nvars = 0;
   if( SCIPconsIsDeleted(cons) || SCIPconsIsModifiable(cons) || SCIPconsIsLocal(cons) || !SCIPconsIsPropagationEnabled(cons) )
      return SCIP_OKAY;

   switch( src )
   {
   case 0:
      n = SCIPgetNVarsLinear(scip, cons);
      cvars = SCIPgetVarsLinear(scip, cons);
      break;
   case 1:
      n = SCIPgetNVarsKnapsack(scip, cons);
      cvars = SCIPgetVarsKnapsack(scip, cons);
      break;
   case 2:
      n = SCIPgetNVarsSetppc(scip, cons);
      cvars = SCIPgetVarsSetppc(scip, cons);
      break;
   case 3:
      n = SCIPgetNVarsLogicor(scip, cons);
      cvars = SCIPgetVarsLogicor(scip, cons);
      break;
   default:
      n = 2;
      break;
   }
   if( n > *size )
   {
      const int newsize = SCIPcalcMemGrowSize(scip, n);
      SCIP_CALL( SCIPreallocBufferArray(scip, vars, newsize) );
      SCIP_CALL( SCIPreallocBufferArray(scip, vals, newsize) );
      *size = newsize;
   }
   switch( src )
   {
   case 0:
   {
      const SCIP_Real* cvals = SCIPgetValsLinear(scip, cons);
      for( v = 0; v < n; ++v )
      {
         (*vars)[v] = cvars[v];
         (*vals)[v] = cvals[v];
      }
      *lhs = SCIPgetLhsLinear(scip, cons);
      *rhs = SCIPgetRhsLinear(scip, cons);
      break;
   }
   case 1:
   {
      /* sum w x <= capacity (cons_knapsack.h:145-177) */
      const SCIP_Longint* weights = SCIPgetWeightsKnapsack(scip, cons);
      for( v = 0; v < n; ++v )
      {
         (*vars)[v] = cvars[v];
         (*vals)[v] = (SCIP_Real)weights[v];
      }
      *lhs = -SCIPinfinity(scip);
      *rhs = (SCIP_Real)SCIPgetCapacityKnapsack(scip, cons);
      break;
   }
   case 2:
   {
      /* sum x == / <= / >= 1 (cons_setppc.h:87-89) */
      const SCIP_SETPPCTYPE type = SCIPgetTypeSetppc(scip, cons);
      for( v = 0; v < n; ++v )
      {
         (*vars)[v] = cvars[v];
         (*vals)[v] = 1.0;
      }
      *lhs = type == SCIP_SETPPCTYPE_PACKING ? -SCIPinfinity(scip) : 1.0;
      *rhs = type == SCIP_SETPPCTYPE_COVERING ? SCIPinfinity(scip) : 1.0;
      break;
   }
   case 3:
      /* sum x >= 1 */
      for( v = 0; v < n; ++v )
      {
         (*vars)[v] = cvars[v];
         (*vals)[v] = 1.0;
      }
      *lhs = 1.0;
      *rhs = SCIPinfinity(scip);
      break;
   default:
      /* lhs <= x + c y <= rhs (cons_varbound.h:140-168) */
      (*vars)[0] = SCIPgetVarVarbound(scip, cons);
      (*vals)[0] = 1.0;
      (*vars)[1] = SCIPgetVbdvarVarbound(scip, cons);
      (*vals)[1] = SCIPgetVbdcoefVarbound(scip, cons);
      *lhs = SCIPgetLhsVarbound(scip, cons);
      *rhs = SCIPgetRhsVarbound(scip, cons);
      break;
   }

   /* active variables: x' = 1 - x, x = a y + c, ... ; the constant moves to the sides (matrix.c: getActiveVariables) */
   SCIP_CALL( SCIPgetProbvarLinearSum(scip, *vars, *vals, &n, *size, &constant, &required) );
   if( required > *size )
   {
      const int newsize = SCIPcalcMemGrowSize(scip, required);
      SCIP_CALL( SCIPreallocBufferArray(scip, vars, newsize) );
      SCIP_CALL( SCIPreallocBufferArray(scip, vals, newsize) );
      *size = newsize;
      SCIP_CALL( SCIPgetProbvarLinearSum(scip, *vars, *vals, &n, *size, &constant, &required) );
      assert(required <= *size);
   }
   for( v = 0; v < n; ++v )
   {
      if( SCIPvarGetProbindex((*vars)[v]) < 0 )
         return SCIP_OKAY;
   }
   if( !SCIPisInfinity(scip, -(*lhs)) )
      *lhs -= constant;
   if( !SCIPisInfinity(scip, *rhs) )
      *rhs -= constant;
   *nvars = n;
   *usable = TRUE;
   return SCIP_OKAY;
}

/** builds the device copy of all usable rows */
static
SCIP_RETCODE buildDeviceCopy(
   SCIP*                 scip,
   SCIP_PROPDATA*        propdata
   )
{
   SCIP_VAR** probvars;
   SCIP_VAR** rowvars;
   SCIP_Real* rowvals;
   gpulin_numerics num;
   uint8_t* vartype;
   int64_t* fill;
   int64_t k;
   int rowsize;
   int nsources;
   int ncols;
   int nrows;
   int pass;
   int src;
   int c;
   int j;
   int rc;

   freeDeviceCopy(scip, propdata);

   nsources = propdata->allrows ? NROWSOURCES : 1;
   propdata->nlinconss = countSourceConss(scip, propdata->allrows, propdata->stablecopy, propdata->sourcetail);
   propdata->nskipped = 0;
   ++propdata->nbuilds;
   if( propdata->nlinconss == 0 )
      return SCIP_OKAY;

   probvars = SCIPgetVars(scip);
   ncols = SCIPgetNVars(scip);
   rowsize = 64;
   SCIP_CALL( SCIPallocBufferArray(scip, &rowvars, rowsize) );
   SCIP_CALL( SCIPallocBufferArray(scip, &rowvals, rowsize) );
   vartype = NULL;
   fill = NULL;

   /* two passes over the constraints of all sources: count, then fill */
   nrows = 0;
   k = 0;
   for( pass = 0; pass < 2; ++pass )
   {
      if( pass == 1 )
      {
         if( nrows == 0 || ncols == 0 )
         {
            SCIPfreeBufferArray(scip, &rowvals);
            SCIPfreeBufferArray(scip, &rowvars);
            return SCIP_OKAY;
         }
         propdata->ncols = ncols;
         propdata->nrows = nrows;
         propdata->nnz = k;
         SCIP_CALL( SCIPallocBlockMemoryArray(scip, &propdata->vars, ncols) );
         SCIP_CALL( SCIPallocBlockMemoryArray(scip, &propdata->lb, ncols) );
         SCIP_CALL( SCIPallocBlockMemoryArray(scip, &propdata->ub, ncols) );
         SCIP_CALL( SCIPallocBlockMemoryArray(scip, &propdata->reflb, ncols) );
         SCIP_CALL( SCIPallocBlockMemoryArray(scip, &propdata->refub, ncols) );
         SCIP_CALL( SCIPallocBlockMemoryArray(scip, &propdata->codes, (ncols + 15) / 16) );
         SCIP_CALL( SCIPallocBlockMemoryArray(scip, &propdata->colptr, ncols + 1) );
         SCIP_CALL( SCIPallocClearBlockMemoryArray(scip, &propdata->rowcons, nrows) );
         SCIP_CALL( SCIPallocBlockMemoryArray(scip, &propdata->lhs, nrows) );
         SCIP_CALL( SCIPallocBlockMemoryArray(scip, &propdata->rhs, nrows) );
         SCIP_CALL( SCIPallocBlockMemoryArray(scip, &propdata->rowptr, nrows + 1) );
         SCIP_CALL( SCIPallocBlockMemoryArray(scip, &propdata->colidx, propdata->nnz) );
         SCIP_CALL( SCIPallocBlockMemoryArray(scip, &propdata->vals, propdata->nnz) );
         SCIP_CALL( SCIPallocBlockMemoryArray(scip, &propdata->colrows, propdata->nnz) );
         SCIP_CALL( SCIPallocBlockMemoryArray(scip, &propdata->colpos, propdata->nnz) );
         SCIP_CALL( SCIPallocBufferArray(scip, &vartype, ncols) );
         SCIP_CALL( SCIPallocBufferArray(scip, &fill, ncols + 1) );
         for( j = 0; j < ncols; ++j )
         {
            assert(SCIPvarGetProbindex(probvars[j]) == j);
            propdata->vars[j] = probvars[j];
            vartype[j] = SCIPvarIsIntegral(probvars[j]) ? GPULIN_VAR_INTEGRAL : GPULIN_VAR_CONTINUOUS;
            propdata->colptr[j] = 0;
         }
         propdata->colptr[ncols] = 0;
         nrows = 0;
         k = 0;
      }
      for( src = 0; src < nsources; ++src )
      {
         SCIP_CONSHDLR* conshdlr = SCIPfindConshdlr(scip, rowsourcenames[src]);
         SCIP_CONS** conss;
         int nconss;

         if( pass == 0 )
            propdata->nrowsof[src] = 0;
         if( conshdlr == NULL )
            continue;
         conss = SCIPconshdlrGetConss(conshdlr);
         nconss = propdata->stablecopy ? SCIPconshdlrGetNConss(conshdlr) : SCIPconshdlrGetNActiveConss(conshdlr);
         for( c = 0; c < nconss; ++c )
         {
            SCIP_Real lhs;
            SCIP_Real rhs;
            SCIP_Bool usable;
            int nvars;
            int v;

            SCIP_CALL( getActiveRow(scip, conss[c], src, &rowvars, &rowvals, &rowsize, &nvars, &lhs, &rhs, &usable) );
            if( !usable )
            {
               if( pass == 0 )
                  ++propdata->nskipped;
               continue;
            }
            if( pass == 1 )
            {
               propdata->rowcons[nrows] = conss[c];
               SCIP_CALL( SCIPcaptureCons(scip, conss[c]) );
               propdata->rowptr[nrows] = k;
               propdata->lhs[nrows] = lhs;
               propdata->rhs[nrows] = rhs;
               for( v = 0; v < nvars; ++v )
               {
                  propdata->colidx[k + v] = SCIPvarGetProbindex(rowvars[v]);
                  propdata->vals[k + v] = rowvals[v];
                  ++propdata->colptr[propdata->colidx[k + v] + 1];
               }
            }
            else
               ++propdata->nrowsof[src];
            k += nvars;
            ++nrows;
         }
      }
   }
   propdata->rowptr[nrows] = k;
   assert(nrows == propdata->nrows && k == propdata->nnz);

   /* column -> rows, for PROPRESPROP */
   for( j = 0; j < ncols; ++j )
      propdata->colptr[j + 1] += propdata->colptr[j];
   for( j = 0; j <= ncols; ++j )
      fill[j] = propdata->colptr[j];
   for( c = 0; c < nrows; ++c )
   {
      for( k = propdata->rowptr[c]; k < propdata->rowptr[c + 1]; ++k )
      {
         int64_t q = fill[propdata->colidx[k]]++;
         propdata->colrows[q] = c;
         propdata->colpos[q] = (int32_t)(k - propdata->rowptr[c]);
      }
   }

   /* tolerances of this SCIP instance */
   num.infinity = SCIPinfinity(scip);
   num.epsilon = SCIPepsilon(scip);
   num.sumepsilon = SCIPsumepsilon(scip);
   num.feastol = SCIPfeastol(scip);
   num.hugeval = SCIPgetHugeValue(scip);
   SCIP_CALL( SCIPgetRealParam(scip, "numerics/boundstreps", &num.boundstreps) );
   SCIP_CALL( SCIPgetRealParam(scip, "constraints/linear/maxeasyactivitydelta", &num.maxeasyactivitydelta) );

   rc = gpulin_create(propdata->device, nrows, ncols, propdata->nnz, propdata->rowptr, propdata->colidx, propdata->vals,
      propdata->lhs, propdata->rhs, vartype, &num, &propdata->gpu);
   /* several devices: SCIP is one process and calls PROPEXEC on one thread (prop.c:646-716), so the copies on the other
    * devices are driven from here -- the same problem on every device, connected in-process */
   while( rc == GPULIN_OK && propdata->npeergpus + 1 < propdata->ndevices )
   {
      rc = gpulin_create(propdata->device + propdata->npeergpus + 1, nrows, ncols, propdata->nnz, propdata->rowptr,
         propdata->colidx, propdata->vals, propdata->lhs, propdata->rhs, vartype, &num, &propdata->peergpu[propdata->npeergpus]);
      if( rc == GPULIN_OK )
         ++propdata->npeergpus;
   }
   /* ranged-row propagation (rangedRowPropagation, cons_linear.c:5715-6696): the device walks a ranged row in the order
    * cons_linear sorts it in (consdataCompVarProp :3191) -- by the global bounds; a column is its variable's probindex */
   if( rc == GPULIN_OK && propdata->rangedrow )
   {
      for( j = 0; j < ncols; ++j )
      {
         propdata->lb[j] = SCIPvarGetLbGlobal(probvars[j]);
         propdata->ub[j] = SCIPvarGetUbGlobal(probvars[j]);
      }
      rc = gpulin_set_rangedrow(propdata->gpu, 1, propdata->lb, propdata->ub, NULL);
      for( j = 0; j < propdata->npeergpus && rc == GPULIN_OK; ++j )
         rc = gpulin_set_rangedrow(propdata->peergpu[j], 1, propdata->lb, propdata->ub, NULL);
   }
   SCIPfreeBufferArray(scip, &fill);
   SCIPfreeBufferArray(scip, &vartype);
   SCIPfreeBufferArray(scip, &rowvals);
   SCIPfreeBufferArray(scip, &rowvars);
   if( rc != GPULIN_OK )
   {
      SCIPerrorMessage("prop_gpulinear: gpulin_create failed (%d): %s\n", rc, gpulin_last_error());
      while( propdata->npeergpus > 0 )
         gpulin_destroy(propdata->peergpu[--propdata->npeergpus]);
      gpulin_destroy(propdata->gpu);
      propdata->gpu = NULL;
      return SCIP_ERROR;
   }

   propdata->logcap = (int64_t)DEFAULT_LOGCAPFAC * ncols + 1024;
   SCIP_CALL( SCIPallocBlockMemoryArray(scip, &propdata->changes, propdata->logcap) );
   SCIP_CALL( SCIPallocBlockMemoryArray(scip, &propdata->clog, propdata->logcap) );
   SCIP_CALL( SCIPallocBlockMemoryArray(scip, &propdata->cside, 3 * propdata->logcap) );
   rc = gpulin_set_change_log(propdata->gpu, propdata->logcap);
   /* (the same log on every device: the devices must take every decision alike, also "the log is full") */
   for( j = 0; j < propdata->npeergpus && rc == GPULIN_OK; ++j )
      rc = gpulin_set_change_log(propdata->peergpu[j], propdata->logcap);
   if( rc != GPULIN_OK )
   {
      SCIPerrorMessage("prop_gpulinear: gpulin_set_change_log failed (%d): %s\n", rc, gpulin_last_error());
      return SCIP_ERROR;
   }
   if( propdata->npeergpus > 0 )
   {
      gpulin_t* group[8];
      group[0] = propdata->gpu;
      for( j = 0; j < propdata->npeergpus; ++j )
         group[j + 1] = propdata->peergpu[j];
      rc = gpulin_group_connect(group, propdata->npeergpus + 1);
      if( rc != GPULIN_OK )
      {
         SCIPerrorMessage("prop_gpulinear: gpulin_group_connect failed (%d): %s\n", rc, gpulin_last_error());
         return SCIP_ERROR;
      }
   }

   /* the reference of the packed bound uploads: the global bounds of now (a column at a node sits at them, or is fixed to
    * one of them, or -- rarely -- has other local bounds: 2 bits per column + an explicit list instead of 16 bytes) */
   for( j = 0; j < ncols; ++j )
   {
      propdata->reflb[j] = SCIPvarGetLbGlobal(propdata->vars[j]);
      propdata->refub[j] = SCIPvarGetUbGlobal(propdata->vars[j]);
   }
   rc = gpulin_set_reference_bounds(propdata->gpu, propdata->reflb, propdata->refub);
   for( j = 0; j < propdata->npeergpus && rc == GPULIN_OK; ++j )
      rc = gpulin_set_reference_bounds(propdata->peergpu[j], propdata->reflb, propdata->refub);
   if( rc != GPULIN_OK )
   {
      SCIPerrorMessage("prop_gpulinear: gpulin_set_reference_bounds failed (%d): %s\n", rc, gpulin_last_error());
      return SCIP_ERROR;
   }

   /* from now on every bound change of a column is noted (cf. consCatchAllEvents, cons_linear.c:717) */
   SCIP_CALL( SCIPallocBlockMemoryArray(scip, &propdata->filterpos, ncols) );
   SCIP_CALL( SCIPallocBlockMemoryArray(scip, &propdata->touched, ncols) );
   SCIP_CALL( SCIPallocBlockMemoryArray(scip, &propdata->istouched, ncols) );
   SCIP_CALL( SCIPallocBlockMemoryArray(scip, &propdata->updidx, ncols) );
   for( j = 0; j < ncols; ++j )
   {
      propdata->filterpos[j] = -1;
      propdata->istouched[j] = FALSE;
   }
   propdata->ntouched = 0;
   propdata->fullsync = TRUE;
   if( propdata->incremental && propdata->eventhdlr != NULL )
   {
      for( j = 0; j < ncols; ++j )
      {
         SCIP_CALL( SCIPcatchVarEvent(scip, propdata->vars[j], SCIP_EVENTTYPE_BOUNDCHANGED, propdata->eventhdlr,
               (SCIP_EVENTDATA*)(size_t)(j + 1), &propdata->filterpos[j]) );
      }
   }

   SCIPverbMessage(scip, SCIP_VERBLEVEL_FULL, NULL,
      "prop_gpulinear: device copy of %d rows (%d linear, %d knapsack, %d setppc, %d logicor, %d varbound; %d skipped), "
      "%d columns, %lld nonzeros\n", nrows, propdata->nrowsof[0], propdata->nrowsof[1], propdata->nrowsof[2],
      propdata->nrowsof[3], propdata->nrowsof[4], propdata->nskipped, ncols, (long long)propdata->nnz);

   return SCIP_OKAY;
}

/** execution method of the event handler: remembers the column (the device copy of its bounds is stale now) */
static
SCIP_DECL_EVENTEXEC(eventExecGpulinear)
{
   SCIP_PROPDATA* propdata = (SCIP_PROPDATA*)SCIPeventhdlrGetData(eventhdlr);
   const int j = (int)(size_t)eventdata - 1;

   (void)scip;
   (void)event;
   assert(propdata != NULL);
   if( propdata->inreplay || propdata->istouched == NULL || j < 0 || j >= propdata->ncols )
      return SCIP_OKAY;
   if( !propdata->istouched[j] )
   {
      propdata->istouched[j] = TRUE;
      propdata->touched[propdata->ntouched++] = j;
   }
   return SCIP_OKAY;
}

/** a bound SCIP holds differs from what the device holds (SCIP adjusted or rejected a replayed change) */
static
void noteMismatch(
   SCIP_PROPDATA*        propdata,
   int                   j
   )
{
   if( propdata->istouched != NULL && !propdata->istouched[j] )
   {
      propdata->istouched[j] = TRUE;
      propdata->touched[propdata->ntouched++] = j;
   }
}

/*
 * Callback methods of propagator
 */

/** destructor of propagator to free user data (called when SCIP is exiting) */
static
SCIP_DECL_PROPFREE(propFreeGpulinear)
{
   SCIP_PROPDATA* propdata;

   propdata = SCIPpropGetData(prop);
   assert(propdata != NULL);
   freeDeviceCopy(scip, propdata);
   SCIPfreeBlockMemory(scip, &propdata);
   SCIPpropSetData(prop, NULL);

   return SCIP_OKAY;
}

/** solving process deinitialization method: the transformed problem goes away, and with it the device copy */
static
SCIP_DECL_PROPEXITSOL(propExitsolGpulinear)
{
   SCIP_PROPDATA* propdata;

   (void)restart;
   propdata = SCIPpropGetData(prop);
   assert(propdata != NULL);
   if( propdata->ncalls > 0 )
   {
      SCIPverbMessage(scip, SCIP_VERBLEVEL_FULL, NULL,
         "prop_gpulinear: %lld calls (%lld device copies built, %lld full bound uploads, %lld single bounds sent), %lld device rounds, "
         "%lld bound changes, %.3f ms on the device\n", (long long)propdata->ncalls, (long long)propdata->nbuilds,
         (long long)propdata->nfullsyncs, (long long)propdata->nupdates,
         (long long)propdata->nrounds, (long long)propdata->nchanges, propdata->devicems);
   }
   freeDeviceCopy(scip, propdata);

   return SCIP_OKAY;
}

/** the change log of the last device call into propdata->changes.  A long log (a root propagation fixes hundreds of thousands
 *  of binaries) comes back compact -- one word per change whose new bound is 0 or 1 instead of 24 bytes, see
 *  gpulin_get_changes_compact; the round of an entry then follows from its position and the per-round counts -- a short one
 *  (the call at a node: a handful of changes) as plain records with one copy and one synchronisation.
 */
#define COMPACTLOG_MINCHANGES  4096
#define COMPACTLOG_MAXROUNDS   1024
static
int fetchChanges(
   SCIP_PROPDATA*        propdata,
   int64_t               nchanges,           /**< accepted changes of the call (gpulin_result) */
   int64_t*              nlog                /**< entries produced */
   )
{
   int64_t roundchg[COMPACTLOG_MAXROUNDS];
   int64_t nside;
   int64_t sum;
   int64_t e;
   int32_t nrounds;
   int rc;
   int r;

   if( nchanges < COMPACTLOG_MINCHANGES || propdata->ncols >= (1 << 29) )
      return gpulin_get_changes(propdata->gpu, propdata->changes, propdata->logcap, nlog);

   rc = gpulin_get_round_stats(propdata->gpu, NULL, NULL, roundchg, COMPACTLOG_MAXROUNDS, &nrounds);
   if( rc != GPULIN_OK )
      return rc;
   sum = 0;
   for( r = 0; r < nrounds; ++r )
      sum += roundchg[r];
   rc = gpulin_get_changes_compact(propdata->gpu, propdata->clog, propdata->logcap, nlog, propdata->cside, propdata->logcap, &nside);
   if( rc != GPULIN_OK )
      return rc;
   if( *nlog > propdata->logcap )
      return GPULIN_OK;                      /* overflow: the caller takes the final bounds instead */
   if( sum != *nlog || nside > propdata->logcap )
      return gpulin_get_changes(propdata->gpu, propdata->changes, propdata->logcap, nlog);   /* (more rounds than the statistics hold) */

   e = 0;
   for( r = 0; r < nrounds; ++r )
   {
      int64_t k;
      for( k = 0; k < roundchg[r]; ++k, ++e )
      {
         const uint32_t w = propdata->clog[e];
         propdata->changes[e].var = (int32_t)(w & 0x1fffffffu);
         propdata->changes[e].round = r;
         propdata->changes[e].is_upper = (int32_t)(w >> 31);
         propdata->changes[e].reserved = 0;
         propdata->changes[e].newbound = ((w >> 29) & 3u) == 1u ? 1.0 : 0.0;
      }
   }
   for( e = 0; e < nside; ++e )
   {
      const uint32_t pos = propdata->cside[3 * e];
      uint64_t bits = (uint64_t)propdata->cside[3 * e + 1] | ((uint64_t)propdata->cside[3 * e + 2] << 32);
      memcpy(&propdata->changes[pos].newbound, &bits, sizeof(double));
   }
   return GPULIN_OK;
}

/** order of the replay: by round (the order that makes every change explainable), inside a round by column and bound.
 *  The device appends the changes of a round in whatever order its atomics land; SCIP's search (inference order,
 *  conflict analysis) should not depend on that */
static
int compareChanges(
   const void*           a,
   const void*           b
   )
{
   const gpulin_change* x = (const gpulin_change*)a;
   const gpulin_change* y = (const gpulin_change*)b;
   if( x->round != y->round )
      return x->round < y->round ? -1 : 1;
   if( x->var != y->var )
      return x->var < y->var ? -1 : 1;
   return x->is_upper - y->is_upper;
}

/** the end of propagateCons (cons_linear.c:7743-7753): rows whose activity bounds lie inside their sides for the bounds of
 *  this node are deleted locally.  Only when the device bounds are SCIP's bounds (no mismatch pending). */
static
SCIP_RETCODE deleteRedundantRows(
   SCIP*                 scip,
   SCIP_PROPDATA*        propdata
   )
{
   int64_t n = 0;
   int64_t i;
   int rc;

   if( !propdata->delredundant || propdata->fullsync || propdata->ntouched > 0 || SCIPinProbing(scip) )
      return SCIP_OKAY;
   if( propdata->redrows == NULL )
      SCIP_CALL( SCIPallocBlockMemoryArray(scip, &propdata->redrows, propdata->nrows) );
   rc = gpulin_get_redundant_rows(propdata->gpu, propdata->redrows, propdata->nrows, &n);
   if( rc != GPULIN_OK )
   {
      SCIPerrorMessage("prop_gpulinear: gpulin_get_redundant_rows failed (%d): %s\n", rc, gpulin_last_error());
      return SCIP_ERROR;
   }
   for( i = 0; i < n && i < propdata->nrows; ++i )
   {
      SCIP_CONS* cons = propdata->rowcons[propdata->redrows[i]];
      if( SCIPconsIsActive(cons) && !SCIPconsIsDeleted(cons) )
      {
         SCIP_CALL( SCIPdelConsLocal(scip, cons) );
         ++propdata->ndelconss;
      }
   }
   return SCIP_OKAY;
}

/** execution method of propagator */
static
SCIP_DECL_PROPEXEC(propExecGpulinear)
{
   SCIP_PROPDATA* propdata;
   gpulin_result res;
   int nsourceconss;
   int64_t nlog;
   int64_t e;
   int ntightened;
   int rc;
   int j;

   (void)proptiming;
   *result = SCIP_DIDNOTRUN;

   propdata = SCIPpropGetData(prop);
   assert(propdata != NULL);

   /* staleness: rebuild when the constraints of the row sources (see countSourceConss) or the variables changed */
   if( deviceCopyIsStale(scip, propdata, &nsourceconss) )
   {
      if( nsourceconss == 0 )
         return SCIP_OKAY;
      SCIP_CALL( buildDeviceCopy(scip, propdata) );
      if( propdata->gpu == NULL )
         return SCIP_OKAY;
   }

   *result = SCIP_DIDNOTFIND;

   /* bounds of the current node: everything after a (re)build or a cutoff, else only what changed since the last call */
   if( propdata->fullsync || !propdata->incremental )
   {
      int nexplicit = 0;
      memset(propdata->codes, 0, sizeof(uint32_t) * (size_t)((propdata->ncols + 15) / 16));
      for( j = 0; j < propdata->ncols; ++j )
      {
         const SCIP_Real l = SCIPvarGetLbLocal(propdata->vars[j]);
         const SCIP_Real u = SCIPvarGetUbLocal(propdata->vars[j]);
         uint32_t code;
         if( l == propdata->reflb[j] && u == propdata->refub[j] ) /*lint !e777*/
            code = 0;
         else if( l == propdata->reflb[j] && u == propdata->reflb[j] ) /*lint !e777*/
            code = 1;
         else if( l == propdata->refub[j] && u == propdata->refub[j] ) /*lint !e777*/
            code = 2;
         else
         {
            code = 3;
            propdata->updidx[nexplicit] = j;
            propdata->lb[nexplicit] = l;
            propdata->ub[nexplicit] = u;
            ++nexplicit;
         }
         propdata->codes[j >> 4] |= code << (2 * (j & 15));
         propdata->istouched[j] = FALSE;
      }
      propdata->ntouched = 0;
      propdata->fullsync = FALSE;
      ++propdata->nfullsyncs;
      rc = gpulin_set_bounds_packed(propdata->gpu, propdata->codes, nexplicit, propdata->updidx, propdata->lb, propdata->ub);
      for( j = 0; j < propdata->npeergpus && rc == GPULIN_OK; ++j )
         rc = gpulin_set_bounds_packed(propdata->peergpu[j], propdata->codes, nexplicit, propdata->updidx, propdata->lb, propdata->ub);
   }
   else
   {
      int n = propdata->ntouched;
      for( j = 0; j < n; ++j )
      {
         const int col = propdata->touched[j];
         propdata->updidx[j] = col;
         propdata->lb[j] = SCIPvarGetLbLocal(propdata->vars[col]);
         propdata->ub[j] = SCIPvarGetUbLocal(propdata->vars[col]);
         propdata->istouched[col] = FALSE;
      }
      propdata->ntouched = 0;
      propdata->nupdates += n;
      rc = gpulin_update_bounds(propdata->gpu, n, propdata->updidx, propdata->lb, propdata->ub);
      for( j = 0; j < propdata->npeergpus && rc == GPULIN_OK; ++j )
         rc = gpulin_update_bounds(propdata->peergpu[j], n, propdata->updidx, propdata->lb, propdata->ub);
   }
   if( rc == GPULIN_OK && propdata->npeergpus == 0 )
      rc = gpulin_propagate(propdata->gpu, propdata->maxrounds < 0 ? 0 : propdata->maxrounds, &res);
   else if( rc == GPULIN_OK )
   {
      /* a collective call: enqueued on every device, then awaited on every device (the devices exchange their candidates
       * among themselves); every device ends with the same bounds, the results are taken from the first */
      gpulin_result peerres;
      rc = gpulin_propagate_async(propdata->gpu, propdata->maxrounds < 0 ? 0 : propdata->maxrounds);
      for( j = 0; j < propdata->npeergpus && rc == GPULIN_OK; ++j )
         rc = gpulin_propagate_async(propdata->peergpu[j], propdata->maxrounds < 0 ? 0 : propdata->maxrounds);
      if( rc == GPULIN_OK )
         rc = gpulin_propagate_wait(propdata->gpu, &res);
      for( j = 0; j < propdata->npeergpus && rc == GPULIN_OK; ++j )
      {
         rc = gpulin_propagate_wait(propdata->peergpu[j], &peerres);
         if( rc == GPULIN_OK && (peerres.status != res.status || peerres.nrounds != res.nrounds || peerres.nchanges != res.nchanges) )
         {
            SCIPerrorMessage("prop_gpulinear: device %d disagrees with device %d\n", propdata->device + j + 1, propdata->device);
            return SCIP_ERROR;
         }
      }
   }
   if( rc != GPULIN_OK )
   {
      SCIPerrorMessage("prop_gpulinear: device propagation failed (%d): %s\n", rc, gpulin_last_error());
      return SCIP_ERROR;
   }
   ++propdata->ncalls;
   propdata->nrounds += res.nrounds;
   propdata->nchanges += res.nchanges;
   propdata->devicems += res.device_ms;

   if( res.status == GPULIN_CUTOFF )
   {
      /* the device bounds moved on a path SCIP never saw: start over from SCIP's bounds next time */
      propdata->fullsync = TRUE;
      *result = SCIP_CUTOFF;
      return SCIP_OKAY;
   }
   if( res.nchanges == 0 )
   {
      SCIP_CALL( deleteRedundantRows(scip, propdata) );
      return SCIP_OKAY;
   }

   ntightened = 0;
   rc = fetchChanges(propdata, res.nchanges, &nlog);
   if( rc != GPULIN_OK )
   {
      SCIPerrorMessage("prop_gpulinear: reading the change log failed (%d): %s\n", rc, gpulin_last_error());
      return SCIP_ERROR;
   }
   if( nlog <= propdata->logcap )
   {
      /* replay in round order: every change is implied by one row and the bounds SCIP knows at that point */
      if( propdata->deterministic && nlog > 1 )
         qsort(propdata->changes, (size_t)nlog, sizeof(gpulin_change), compareChanges);
      for( e = 0; e < nlog; ++e )
      {
         const gpulin_change* chg = &propdata->changes[e];
         SCIP_VAR* var = propdata->vars[chg->var];
         SCIP_Bool infeasible;
         SCIP_Bool tightened;
         SCIP_RETCODE retcode;

         propdata->inreplay = TRUE;      /* the device already holds this bound */
         if( chg->is_upper )
            retcode = SCIPinferVarUbProp(scip, var, chg->newbound, prop, 1, TRUE, &infeasible, &tightened);
         else
            retcode = SCIPinferVarLbProp(scip, var, chg->newbound, prop, 0, TRUE, &infeasible, &tightened);
         propdata->inreplay = FALSE;
         SCIP_CALL( retcode );
         if( infeasible )
         {
            propdata->fullsync = TRUE;
            *result = SCIP_CUTOFF;
            return SCIP_OKAY;
         }
         if( tightened )
            ++ntightened;
         /* SCIP may have adjusted or declined the value: then the device copy of this column is stale */
         if( (chg->is_upper ? SCIPvarGetUbLocal(var) : SCIPvarGetLbLocal(var)) != chg->newbound )
            noteMismatch(propdata, chg->var);
      }
   }
   else
   {
      /* the log overflowed: hand over the final bounds (no valid replay order: PROPRESPROP may fail for them) */
      propdata->fullsync = TRUE;
      rc = gpulin_get_bounds(propdata->gpu, propdata->lb, propdata->ub);
      if( rc != GPULIN_OK )
      {
         SCIPerrorMessage("prop_gpulinear: gpulin_get_bounds failed (%d): %s\n", rc, gpulin_last_error());
         return SCIP_ERROR;
      }
      for( j = 0; j < propdata->ncols; ++j )
      {
         SCIP_Bool infeasible;
         SCIP_Bool tightened;

         if( propdata->lb[j] > SCIPvarGetLbLocal(propdata->vars[j]) )
         {
            SCIP_CALL( SCIPinferVarLbProp(scip, propdata->vars[j], propdata->lb[j], prop, 0, TRUE, &infeasible, &tightened) );
            if( infeasible )
            {
               *result = SCIP_CUTOFF;
               return SCIP_OKAY;
            }
            ntightened += tightened ? 1 : 0;
         }
         if( propdata->ub[j] < SCIPvarGetUbLocal(propdata->vars[j]) )
         {
            SCIP_CALL( SCIPinferVarUbProp(scip, propdata->vars[j], propdata->ub[j], prop, 1, TRUE, &infeasible, &tightened) );
            if( infeasible )
            {
               *result = SCIP_CUTOFF;
               return SCIP_OKAY;
            }
            ntightened += tightened ? 1 : 0;
         }
      }
   }
   if( ntightened > 0 )
      *result = SCIP_REDUCEDDOM;

   SCIP_CALL( deleteRedundantRows(scip, propdata) );

   return SCIP_OKAY;
}

/** propagation conflict resolving method of propagator: finds a row of the column of infervar whose residual activity
 *  at bdchgidx implies the inferred bound and reports the bounds of its other variables -- the simple variant of
 *  cons_linear's resolvePropagation (cons_linear.c:4951-4971 via addConflictBounds :4690) */
static
SCIP_DECL_PROPRESPROP(propRespropGpulinear)
{
   SCIP_PROPDATA* propdata;
   int64_t q;
   int j;

   (void)inferinfo;
   *result = SCIP_DIDNOTFIND;
   propdata = SCIPpropGetData(prop);
   assert(propdata != NULL);
   if( propdata->gpu == NULL )
      return SCIP_OKAY;
   j = SCIPvarGetProbindex(infervar);
   if( j < 0 || j >= propdata->ncols || propdata->vars[j] != infervar )
      return SCIP_OKAY;

   for( q = propdata->colptr[j]; q < propdata->colptr[j + 1]; ++q )
   {
      const int r = propdata->colrows[q];
      const int64_t beg = propdata->rowptr[r];
      const int64_t end = propdata->rowptr[r + 1];
      const SCIP_Real a = propdata->vals[beg + propdata->colpos[q]];
      /* an upper bound of x (a > 0) or a lower bound (a < 0) comes from  rhs - minresidual ; the other two cases from
       * lhs - maxresidual */
      const SCIP_Bool userhs = (boundtype == SCIP_BOUNDTYPE_UPPER) == (a > 0.0);
      const SCIP_Real side = userhs ? propdata->rhs[r] : propdata->lhs[r];
      SCIP_Real resact = 0.0;
      SCIP_Bool finite = TRUE;
      SCIP_Real implied;
      int64_t k;

      if( SCIPisInfinity(scip, userhs ? side : -side) )
         continue;
      for( k = beg; k < end && finite; ++k )
      {
         SCIP_VAR* var = propdata->vars[propdata->colidx[k]];
         SCIP_Real bnd;
         if( k == beg + propdata->colpos[q] )
            continue;
         /* minimal residual activity: lb for positive, ub for negative coefficients; maximal: the other way round */
         if( (propdata->vals[k] > 0.0) == userhs )
            bnd = SCIPgetVarLbAtIndex(scip, var, bdchgidx, FALSE);
         else
            bnd = SCIPgetVarUbAtIndex(scip, var, bdchgidx, FALSE);
         if( SCIPisInfinity(scip, REALABS(bnd)) )
            finite = FALSE;
         else
            resact += propdata->vals[k] * bnd;
      }
      if( !finite )
         continue;
      implied = (side - resact) / a;
      if( boundtype == SCIP_BOUNDTYPE_UPPER )
      {
         if( SCIPvarIsIntegral(infervar) )
            implied = SCIPfeasFloor(scip, implied);
         if( !SCIPisFeasLE(scip, implied, relaxedbd) )
            continue;
      }
      else
      {
         if( SCIPvarIsIntegral(infervar) )
            implied = SCIPfeasCeil(scip, implied);
         if( !SCIPisFeasGE(scip, implied, relaxedbd) )
            continue;
      }
      /* this row explains the bound: report the bounds of all its other variables */
      for( k = beg; k < end; ++k )
      {
         SCIP_VAR* var = propdata->vars[propdata->colidx[k]];
         if( k == beg + propdata->colpos[q] )
            continue;
         if( (propdata->vals[k] > 0.0) == userhs )
            SCIP_CALL( SCIPaddConflictLb(scip, var, bdchgidx) );
         else
            SCIP_CALL( SCIPaddConflictUb(scip, var, bdchgidx) );
      }
      *result = SCIP_SUCCESS;
      return SCIP_OKAY;
   }

   return SCIP_OKAY;
}

/*
 * propagator specific interface methods
 */

/** creates the gpulinear propagator and includes it in SCIP */
SCIP_RETCODE SCIPincludePropGpulinear(
   SCIP*                 scip                /**< SCIP data structure */
   )
{
   SCIP_PROPDATA* propdata;
   SCIP_PROP* prop;

   SCIP_CALL( SCIPallocBlockMemory(scip, &propdata) );
   memset(propdata, 0, sizeof(*propdata));
   propdata->nlinconss = -1;

   prop = NULL;
   SCIP_CALL( SCIPincludePropBasic(scip, &prop, PROP_NAME, PROP_DESC, PROP_PRIORITY, PROP_FREQ, PROP_DELAY, PROP_TIMING,
         propExecGpulinear, propdata) );
   assert(prop != NULL);

   SCIP_CALL( SCIPincludeEventhdlrBasic(scip, &propdata->eventhdlr, EVENTHDLR_NAME, EVENTHDLR_DESC, eventExecGpulinear,
         (SCIP_EVENTHDLRDATA*)propdata) );

   /* PROPCOPY stays NULL: sub-SCIPs and concurrent copies do not get a GPU propagator */
   SCIP_CALL( SCIPsetPropFree(scip, prop, propFreeGpulinear) );
   SCIP_CALL( SCIPsetPropExitsol(scip, prop, propExitsolGpulinear) );
   SCIP_CALL( SCIPsetPropResprop(scip, prop, propRespropGpulinear) );

   SCIP_CALL( SCIPaddIntParam(scip, "propagating/" PROP_NAME "/maxdevicerounds",
         "maximal number of propagation rounds per call on the device (-1: to the fixpoint)",
         &propdata->maxrounds, FALSE, DEFAULT_MAXROUNDS, -1, INT_MAX, NULL, NULL) );
   SCIP_CALL( SCIPaddBoolParam(scip, "propagating/" PROP_NAME "/incremental",
         "send only the bounds that changed since the last call (bound change events) instead of all bounds",
         &propdata->incremental, FALSE, DEFAULT_INCREMENTAL, NULL, NULL) );
   SCIP_CALL( SCIPaddBoolParam(scip, "propagating/" PROP_NAME "/allrows",
         "also propagate the linear rows behind knapsack, setppc, logicor and varbound constraints (cf. matrix.c)",
         &propdata->allrows, FALSE, DEFAULT_ALLROWS, NULL, NULL) );
   SCIP_CALL( SCIPaddBoolParam(scip, "propagating/" PROP_NAME "/rangedrow",
         "should the device run the ranged-row (gcd) propagation of equations and ranged rows as well (the counterpart of constraints/linear/rangedrowpropagation, without artificial constraints)?",
         &propdata->rangedrow, FALSE, FALSE, NULL, NULL) );
   SCIP_CALL( SCIPaddBoolParam(scip, "propagating/" PROP_NAME "/delredundant",
         "delete rows locally that the device finds redundant for the node's bounds (the end of propagateCons; cons_linear does this itself for its own rows)",
         &propdata->delredundant, FALSE, DEFAULT_DELREDUNDANT, NULL, NULL) );
   SCIP_CALL( SCIPaddBoolParam(scip, "propagating/" PROP_NAME "/deterministic",
         "replay the bound changes of a device round sorted by variable (the device logs them in the order its atomics land)",
         &propdata->deterministic, FALSE, DEFAULT_DETERMINISTIC, NULL, NULL) );
   SCIP_CALL( SCIPaddBoolParam(scip, "propagating/" PROP_NAME "/stablecopy",
         "keep one device copy of all existing global rows across the tree (FALSE: of the rows active at the node of the build, rebuilt whenever that number changes)",
         &propdata->stablecopy, FALSE, DEFAULT_STABLECOPY, NULL, NULL) );
   SCIP_CALL( SCIPaddIntParam(scip, "propagating/" PROP_NAME "/ndevices",
         "number of CUDA devices (device, device + 1, ...) that share the dense propagation rounds",
         &propdata->ndevices, TRUE, 1, 1, 8, NULL, NULL) );
   SCIP_CALL( SCIPaddIntParam(scip, "propagating/" PROP_NAME "/device",
         "CUDA device ordinal",
         &propdata->device, TRUE, DEFAULT_DEVICE, 0, 1023, NULL, NULL) );

   return SCIP_OKAY;
}

/** rows on the device by source (0 linear, 1 knapsack, 2 setppc, 3 logicor, 4 varbound) */
int SCIPgetNRowsGpulinear(
   SCIP*                 scip,               /**< SCIP data structure */
   int                   source              /**< row source */
   )
{
   SCIP_PROP* prop = SCIPfindProp(scip, PROP_NAME);
   SCIP_PROPDATA* propdata;

   if( prop == NULL || source < 0 || source >= NROWSOURCES )
      return -1;
   propdata = SCIPpropGetData(prop);
   if( propdata == NULL || propdata->gpu == NULL )
      return -1;
   return propdata->nrowsof[source];
}

/** device copies built so far (a rebuild follows every change of the number of source constraints) */
SCIP_Longint SCIPgetNBuildsGpulinear(
   SCIP*                 scip                /**< SCIP data structure */
   )
{
   SCIP_PROP* prop = SCIPfindProp(scip, PROP_NAME);
   SCIP_PROPDATA* propdata;

   if( prop == NULL )
      return -1;
   propdata = SCIPpropGetData(prop);
   return propdata == NULL ? -1 : propdata->nbuilds;
}

/** a batch of independent probes on the current node, see prop_gpulinear.h */
/** the node first: device copy up to date, node bounds on the device, node fixpoint replayed into SCIP; a second pass
 *  takes the values SCIP adjusted during the replay back to the device; then the columns of the probed variables */
static
SCIP_RETCODE prepareProbeBatch(
   SCIP*                 scip,
   SCIP_PROP*            prop,
   SCIP_PROPDATA*        propdata,
   int                   nprobes,
   SCIP_VAR**            vars,
   SCIP_Bool*            nodecutoff,
   int32_t*              cols
   )
{
   SCIP_RESULT result;
   int iter;
   int i;

   if( nodecutoff != NULL )
      *nodecutoff = FALSE;
   for( iter = 0; iter < 4; ++iter )
   {
      SCIP_CALL( propExecGpulinear(scip, prop, SCIP_PROPTIMING_BEFORELP, &result) );
      if( result == SCIP_CUTOFF )
      {
         if( nodecutoff != NULL )
            *nodecutoff = TRUE;
         return SCIP_OKAY;
      }
      if( result != SCIP_REDUCEDDOM && propdata->ntouched == 0 && !propdata->fullsync )
         break;
   }
   if( propdata->gpu == NULL )
      return SCIP_OKAY;
   for( i = 0; i < nprobes; ++i )
   {
      cols[i] = SCIPvarGetProbindex(vars[i]);
      if( cols[i] < 0 || cols[i] >= propdata->ncols || propdata->vars[cols[i]] != vars[i] )
      {
         SCIPerrorMessage("prop_gpulinear: probe %d: <%s> is not an active problem variable\n", i, SCIPvarGetName(vars[i]));
         return SCIP_INVALIDDATA;
      }
   }
   return SCIP_OKAY;
}

SCIP_RETCODE SCIPprobeBatchGpulinear(
   SCIP*                 scip,               /**< SCIP data structure */
   int                   nprobes,            /**< number of probes */
   SCIP_VAR**            vars,               /**< active problem variable of each probe */
   SCIP_Real*            lbs,                /**< lower bound of the variable in its probe */
   SCIP_Real*            ubs,                /**< upper bound of the variable in its probe */
   SCIP_Bool*            nodecutoff,         /**< the node itself is infeasible: no probe was run */
   SCIP_Bool*            cutoff,             /**< per probe: the probe is infeasible */
   int*                  nrounds,            /**< per probe: propagation rounds */
   SCIP_Longint*         nchgbds             /**< per probe: bound changes found */
   )
{
   SCIP_PROP* prop = SCIPfindProp(scip, PROP_NAME);
   SCIP_PROPDATA* propdata;
   SCIP_Bool nodecut = FALSE;
   SCIP_RETCODE retcode;
   int32_t* cols;
   int32_t* status;
   int32_t* rounds;
   int64_t* nchanges;
   int rc;
   int i;

   if( prop == NULL )
   {
      SCIPerrorMessage("prop_gpulinear is not included\n");
      return SCIP_PLUGINNOTFOUND;
   }
   propdata = SCIPpropGetData(prop);
   assert(propdata != NULL);
   for( i = 0; i < nprobes; ++i )
   {
      if( cutoff != NULL )
         cutoff[i] = FALSE;
      if( nrounds != NULL )
         nrounds[i] = 0;
      if( nchgbds != NULL )
         nchgbds[i] = 0;
   }
   SCIP_CALL( SCIPallocBufferArray(scip, &cols, nprobes + 1) );
   retcode = prepareProbeBatch(scip, prop, propdata, nprobes, vars, &nodecut, cols);
   if( nodecutoff != NULL )
      *nodecutoff = nodecut;
   if( retcode != SCIP_OKAY || nodecut || propdata->gpu == NULL || nprobes == 0 )
   {
      SCIPfreeBufferArray(scip, &cols);
      return retcode;
   }
   SCIP_CALL( SCIPallocBufferArray(scip, &status, nprobes) );
   SCIP_CALL( SCIPallocBufferArray(scip, &rounds, nprobes) );
   SCIP_CALL( SCIPallocBufferArray(scip, &nchanges, nprobes) );
   rc = gpulin_probe_batch(propdata->gpu, 32, nprobes, cols, lbs, ubs, propdata->maxrounds < 0 ? 0 : propdata->maxrounds,
      status, rounds, nchanges);
   if( rc == GPULIN_OK )
   {
      for( i = 0; i < nprobes; ++i )
      {
         if( cutoff != NULL )
            cutoff[i] = (status[i] == GPULIN_CUTOFF);
         if( nrounds != NULL )
            nrounds[i] = rounds[i];
         if( nchgbds != NULL )
            nchgbds[i] = nchanges[i];
      }
   }
   SCIPfreeBufferArray(scip, &nchanges);
   SCIPfreeBufferArray(scip, &rounds);
   SCIPfreeBufferArray(scip, &status);
   SCIPfreeBufferArray(scip, &cols);
   if( rc != GPULIN_OK )
   {
      SCIPerrorMessage("SCIPprobeBatchGpulinear failed (%d): %s\n", rc, gpulin_last_error());
      return SCIP_ERROR;
   }
   return SCIP_OKAY;
}

SCIP_RETCODE SCIPprobeBatchBoundsGpulinear(
   SCIP*                 scip,               /**< SCIP data structure */
   int                   nprobes,            /**< number of probes */
   SCIP_VAR**            vars,               /**< active problem variable of each probe */
   SCIP_Real*            lbs,                /**< lower bound of the variable in its probe */
   SCIP_Real*            ubs,                /**< upper bound of the variable in its probe */
   SCIP_Bool*            nodecutoff,         /**< the node itself is infeasible: no probe was run */
   SCIP_Bool*            cutoff,             /**< per probe: the probe is infeasible */
   int*                  chgbeg,             /**< per probe: first entry of its bound changes (nprobes + 1 entries) */
   SCIP_VAR**            chgvars,            /**< variable of every bound change */
   SCIP_BOUNDTYPE*       chgtypes,           /**< bound type of every bound change */
   SCIP_Real*            chgbounds,          /**< new bound of every bound change */
   int                   maxchgs,            /**< capacity of the three arrays */
   int*                  nchgs               /**< bound changes produced */
   )
{
   SCIP_PROP* prop = SCIPfindProp(scip, PROP_NAME);
   SCIP_PROPDATA* propdata;
   SCIP_Bool nodecut = FALSE;
   SCIP_RETCODE retcode;
   gpulin_change* chg;
   int32_t* cols;
   int32_t* status;
   int64_t* beg;
   int64_t nchg = 0;
   int rc;
   int i;

   if( prop == NULL )
   {
      SCIPerrorMessage("prop_gpulinear is not included\n");
      return SCIP_PLUGINNOTFOUND;
   }
   if( chgbeg == NULL || nchgs == NULL || maxchgs < 0 )
      return SCIP_INVALIDDATA;
   propdata = SCIPpropGetData(prop);
   assert(propdata != NULL);
   *nchgs = 0;
   for( i = 0; i <= nprobes; ++i )
      chgbeg[i] = 0;
   for( i = 0; i < nprobes && cutoff != NULL; ++i )
      cutoff[i] = FALSE;
   SCIP_CALL( SCIPallocBufferArray(scip, &cols, nprobes + 1) );
   retcode = prepareProbeBatch(scip, prop, propdata, nprobes, vars, &nodecut, cols);
   if( nodecutoff != NULL )
      *nodecutoff = nodecut;
   if( retcode != SCIP_OKAY || nodecut || propdata->gpu == NULL || nprobes == 0 )
   {
      SCIPfreeBufferArray(scip, &cols);
      return retcode;
   }
   SCIP_CALL( SCIPallocBufferArray(scip, &status, nprobes) );
   SCIP_CALL( SCIPallocBufferArray(scip, &beg, nprobes + 1) );
   SCIP_CALL( SCIPallocBufferArray(scip, &chg, maxchgs + 1) );
   rc = gpulin_probe_batch_changes(propdata->gpu, 32, nprobes, cols, lbs, ubs, propdata->maxrounds < 0 ? 0 : propdata->maxrounds,
      status, NULL, NULL, beg, chg, maxchgs, &nchg);
   if( rc == GPULIN_OK || (rc == GPULIN_ERR_ARG && nchg > maxchgs) )
   {
      for( i = 0; i < nprobes; ++i )
      {
         if( cutoff != NULL )
            cutoff[i] = (status[i] == GPULIN_CUTOFF);
         chgbeg[i] = (int)beg[i];
      }
      chgbeg[nprobes] = (int)beg[nprobes];
      *nchgs = (int)nchg;
      if( rc == GPULIN_OK )
      {
         for( i = 0; i < (int)nchg; ++i )
         {
            chgvars[i] = propdata->vars[chg[i].var];
            chgtypes[i] = chg[i].is_upper ? SCIP_BOUNDTYPE_UPPER : SCIP_BOUNDTYPE_LOWER;
            chgbounds[i] = chg[i].newbound;
         }
      }
      rc = GPULIN_OK;
   }
   SCIPfreeBufferArray(scip, &chg);
   SCIPfreeBufferArray(scip, &beg);
   SCIPfreeBufferArray(scip, &status);
   SCIPfreeBufferArray(scip, &cols);
   if( rc != GPULIN_OK )
   {
      SCIPerrorMessage("SCIPprobeBatchBoundsGpulinear failed (%d): %s\n", rc, gpulin_last_error());
      return SCIP_ERROR;
   }
   return SCIP_OKAY;
}
