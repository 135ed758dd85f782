/* finitediff.h -- forwards to cnavier_dropin.h, which declares the reference's include/finitediff.h interface
 * as served by libcnavier_dropin.so (B200 path). */
#ifndef CNV_FWD_FINITEDIFF_H_INCLUDED
#define CNV_FWD_FINITEDIFF_H_INCLUDED
#include "cnavier_dropin.h"
#endif
