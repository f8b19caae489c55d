/* Forwarding header: lets a source file written against the reference's
 * cpp/grid-algorithm/inc/kernels.h (e.g. the reference's own src/grid.cu, which includes
 * "../inc/kernels.h") compile against libcge instead.  Everything lives in cge_compat.h. */
#include "../../cge_compat.h"
