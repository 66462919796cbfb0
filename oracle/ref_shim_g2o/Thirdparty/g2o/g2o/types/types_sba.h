// -*- C++ -*-
// TEST INFRASTRUCTURE (oracle/ref_shim): the landmark vertex of g2o/types/types_sba.h (a 3-vector with additive update).
#pragma once
#include "../core/base_vertex.h"
namespace g2o {
class VertexSBAPointXYZ : public BaseVertex<3, Eigen::Vector3d> {
 public:
  bool read(std::istream&) override { return false; }
  bool write(std::ostream&) const override { return false; }
  void setToOriginImpl() override { _estimate.setZero(); }
  void oplusImpl(const double* update) override { for (int i = 0; i < 3; ++i) _estimate(i) += update[i]; }
};
}  // namespace g2o
