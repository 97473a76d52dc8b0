// reconstruction kernel instantiations (tile + cooperative) for n_dims = 3, high-order stencil degree 1 (order 2).
#include "recon_inst.cuh"
namespace zfvm {
ZFVM_DEFINE_RECON(3, 1, 9, 6)
}
