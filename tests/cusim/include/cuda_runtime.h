// stand-in for <cuda_runtime.h> in the host-side kernel emulation build (tests/cusim): test infrastructure
#pragma once
#include "../cusim.h"
