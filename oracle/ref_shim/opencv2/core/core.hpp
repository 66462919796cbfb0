// Test infrastructure (oracle/ref_shim): the reference headers include OpenCV, the pinned sources never use it.
#pragma once
