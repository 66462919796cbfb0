// Test infrastructure (oracle/ref_shim): boost::function as the reference's include/Pose3utils.h:26 uses it.
#pragma once
#include <functional>
namespace boost { template <class Sig> using function = std::function<Sig>; }
