// fused-step kernel instantiations for CTAs of 160 threads (see variants.h)
#include "variants.h"

namespace swalbe {
SW_DEFINE_VARIANT(160, 4, 3)
}
