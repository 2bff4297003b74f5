#ifndef POLEE_TF_STUB_FRAMEWORK_OP_KERNEL_H
#define POLEE_TF_STUB_FRAMEWORK_OP_KERNEL_H
#include "tf_stub_core.h"
#endif
