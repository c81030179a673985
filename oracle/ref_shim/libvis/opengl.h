// Build shim used ONLY by oracle/build_ref.sh: the reference's CUDA headers
// include <libvis/opengl.h> (GLEW) without using any GL symbol in the BA
// kernels.  GLEW is not installed in this image, so this 3-line stand-in
// (found first on the include path) supplies what the real header pulls in
// transitively.  It is not a copy of the reference header.
#pragma once
#include "libvis/logging.h"
#include "libvis/libvis.h"
