// reconstruction kernel instantiations (tile + cooperative) for n_dims = 2, high-order stencil degree 1 (order 2).
#include "recon_inst.cuh"
namespace zfvm {
ZFVM_DEFINE_RECON(2, 1, 6, 4)
}
