// Definitions of the opaque host handles of include/zfvm.h.
#pragma once
#include "zfvm_host.hpp"

struct zfvm_grid {
  zfvm::HostGrid g;
};

struct zfvm_stencils {
  zfvm::HostStencils s;
};
