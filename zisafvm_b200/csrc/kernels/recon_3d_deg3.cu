// reconstruction kernel instantiations (tile + cooperative) for n_dims = 3, high-order stencil degree 3 (order 4).
#include "recon_inst.cuh"
namespace zfvm {
ZFVM_DEFINE_RECON(3, 3, 57, 6)
}
