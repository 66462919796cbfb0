// Test infrastructure (oracle/ref_shim_tracking): see Frame.h beside this file.
#pragma once
#include "Frame.h"
