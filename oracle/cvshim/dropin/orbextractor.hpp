// TEST INFRASTRUCTURE ONLY.  What INTEGRATION.md section 1 tells a maintainer to do with the reference's
// include/orbextractor.hpp: it becomes a one-line forwarder to the liborbx-backed class of the same name.  With this
// directory ahead of the reference's include/ on the include path, the reference's own src/orbframe.cpp -- UNMODIFIED --
// constructs, calls and reads the drop-in OrbExtractor (oracle/_ref/libdropinref.so, tests/test_gpu_dropin_frame.py).
#pragma once
#include "orbextractor_b200.hpp"
