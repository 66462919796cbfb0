// Test infrastructure (oracle/ref_shim): src/G2oTypes.cc includes this header and uses nothing from it.
#pragma once
