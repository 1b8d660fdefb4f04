/* oracle/shim/proj.h — TEST INFRASTRUCTURE ONLY.  util/Transformation.h includes <proj.h> for the PJ
 * handle type of its Proj4Transform; the tiler hot path only ever uses the identity transform, PROJ is not
 * installed in this image, and nothing here calls into it. */
#pragma once
typedef struct PJconsts PJ;
typedef struct pj_ctx PJ_CONTEXT;
