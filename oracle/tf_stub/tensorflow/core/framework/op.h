// Minimal stand-in for the TensorFlow C++ op API -- TEST INFRASTRUCTURE ONLY.
// Purpose: compile the reference's src/tensorflow_ext/hsb_ops.cpp UNMODIFIED (from where it lies
// under /root/reference) into oracle/_ref/ so its three CPU OpKernels can be executed here as the
// ground truth for the HSB / InvHSB / InvHSBGrad parity tests.  Only what that one file touches
// is provided.  Written from the public TF op-writing guide's API shape; no TF source was used.
#ifndef POLEE_TF_STUB_OP_H
#define POLEE_TF_STUB_OP_H
#include "tf_stub_core.h"
#endif
