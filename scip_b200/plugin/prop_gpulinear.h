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

#ifdef __cplusplus
}
#endif

#endif
