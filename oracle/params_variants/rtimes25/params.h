/* TEST INFRASTRUCTURE ONLY (oracle). Parameter override for a variant build of the UNMODIFIED reference
 * (oracle/Makefile, _ref/libgpisref_rtimes25.so): this directory comes first on the include path, pulls in the
 * reference's own params.h and changes one macro — the training-ball radius multiple of BASELINE configs[4]
 * ("max points per leaf raised": the reference has no direct knob, N follows from Rtimes). */
#include_next "params.h"
#undef GPISMAP3_RTIMES
#define GPISMAP3_RTIMES 2.5
