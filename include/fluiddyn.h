/* fluiddyn.h -- forwards to cnavier_dropin.h, which declares the reference's include/fluiddyn.h interface
 * as served by libcnavier_dropin.so (B200 path). */
#ifndef CNV_FWD_FLUIDDYN_H_INCLUDED
#define CNV_FWD_FLUIDDYN_H_INCLUDED
#include "cnavier_dropin.h"
#endif
