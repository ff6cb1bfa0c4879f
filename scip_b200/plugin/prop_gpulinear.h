/* prop_gpulinear.h -- B200 linear bound propagation as an ordinary SCIP propagator plugin.
 *
 * Replaces the tightenBounds / activity path of the linear constraint handler (cons_linear.c:6980-7154, :5380-5653,
 * :6700-6974) by the CUDA library libgpulin.so (include/gpulin.h).  Registered like any other propagator:
 *
 *    SCIP_CALL( SCIPincludeDefaultPlugins(scip) );
 *    SCIP_CALL( SCIPincludePropGpulinear(scip) );
 *    SCIP_CALL( SCIPsetIntParam(scip, "constraints/linear/tightenboundsfreq", -1) );   // switch the CPU path off
 */
#ifndef __SCIP_PROP_GPULINEAR_H__
#define __SCIP_PROP_GPULINEAR_H__

#include "scip/def.h"
#include "scip/type_retcode.h"
#include "scip/type_scip.h"
#include "scip/type_var.h"
#include "scip/type_lp.h"

#ifdef __cplusplus
extern "C" {
#endif

/** creates the gpulinear propagator and includes it in SCIP (SCIPincludePropBasic, scip_prop.h:107) */
SCIP_RETCODE SCIPincludePropGpulinear(
   SCIP*                 scip                /**< SCIP data structure */
   );

/** rows on the device by source (0 linear, 1 knapsack, 2 setppc, 3 logicor, 4 varbound; cf. SCIPmatrixGetNRows,
 *  pub_matrix.h); -1 if there is no device copy or the propagator is not included */
int SCIPgetNRowsGpulinear(
   SCIP*                 scip,               /**< SCIP data structure */
   int                   source              /**< row source */
   );

/** device copies built so far (a rebuild follows every change of the number of source constraints); -1 if the
 *  propagator is not included */
SCIP_Longint SCIPgetNBuildsGpulinear(
   SCIP*                 scip                /**< SCIP data structure */
   );

/** a batch of independent probes on the current node -- the cycle SCIPstartProbing / SCIPchgVarLb/UbProbing /
 *  SCIPpropagateProbing / SCIPbacktrackProbing (scip_probing.c:120,302,346,581,226) that SCIPapplyProbingVar
 *  (prop_probing.c:1254-1279) runs once per candidate, for all candidates at once: probe i sets vars[i] to
 *  [lbs[i], ubs[i]] on top of the node's bounds and propagates the device rows to the fixpoint (gpulin_probe_batch: one
 *  launch per probe, many probes in flight).  The node itself is propagated first (like a call of the propagator; its
 *  reductions are applied, *nodecutoff reports an infeasible node).  Outputs per probe (each array may be NULL):
 *  cutoff, propagation rounds, number of bound changes.  Stage SOLVING, at a node or in probing. */
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
   );

/** the same batch, and what every probe implied -- the proplbs / propubs that SCIPapplyProbingVar (prop_probing.c:1203-1303)
 *  returns per candidate, in sparse form: the bound changes of probe i are the entries chgbeg[i] .. chgbeg[i+1] - 1 of
 *  (chgvars, chgtypes, chgbounds), in the order they were found (a variable may appear more than once: the last entry is
 *  its bound at the probe's fixpoint; the probed variable itself is not listed unless propagation tightened it further).
 *  chgbeg needs nprobes + 1 entries.  If the batch produced more than maxchgs entries, *nchgs says how many are needed and
 *  only chgbeg is valid. */
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
   );

#ifdef __cplusplus
}
#endif

#endif
