// fused-step kernel instantiations for CTAs of 128 threads (see variants.h)
#include "variants.h"

namespace swalbe {
SW_DEFINE_VARIANT(128, 5, 3)
}
