// reconstruction kernel instantiations (tile + cooperative) for n_dims = 2, high-order stencil degree 4 (order 5).
#include "recon_inst.cuh"
namespace zfvm {
ZFVM_DEFINE_RECON(2, 4, 28, 3)
}
