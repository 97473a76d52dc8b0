// reconstruction kernel instantiations (tile + cooperative) for n_dims = 3, high-order stencil degree 2 (order 3).
#include "recon_inst.cuh"
namespace zfvm {
ZFVM_DEFINE_RECON(3, 2, 18, 4)
}
