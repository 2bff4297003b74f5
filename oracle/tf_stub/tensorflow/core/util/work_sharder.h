#ifndef POLEE_TF_STUB_UTIL_WORK_SHARDER_H
#define POLEE_TF_STUB_UTIL_WORK_SHARDER_H
#include "tf_stub_core.h"
#endif
