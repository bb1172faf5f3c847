// fused-step kernel instantiations for CTAs of 224 threads (see variants.h)
#include "variants.h"

namespace swalbe {
SW_DEFINE_VARIANT(224, 3, 2)
}
