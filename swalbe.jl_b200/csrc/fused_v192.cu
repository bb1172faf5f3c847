// fused-step kernel instantiations for CTAs of 192 threads (see variants.h)
#include "variants.h"

namespace swalbe {
SW_DEFINE_VARIANT(192, 3, 2)
}
