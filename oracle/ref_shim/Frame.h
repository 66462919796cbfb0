// Test infrastructure (oracle/ref_shim): see KeyFrame.h beside this file.
#pragma once
#include "KeyFrame.h"
