// fused-step kernel instantiations for CTAs of 256 threads (see variants.h)
#include "variants.h"

namespace swalbe {
SW_DEFINE_VARIANT(256, 2, 2)
}
