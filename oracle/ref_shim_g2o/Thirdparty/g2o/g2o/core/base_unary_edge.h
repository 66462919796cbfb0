// -*- C++ -*-
// TEST INFRASTRUCTURE (oracle/ref_shim): see base_vertex.h beside this file.
#pragma once
#include "base_vertex.h"
namespace g2o {
template <int D, class E, class VertexXi> class BaseUnaryEdge : public BaseEdge<D, E> {
 public:
  typedef Eigen::Matrix<double, D, VertexXi::Dimension> JacobianXiOplusType;
  BaseUnaryEdge() { this->_vertices.resize(1, nullptr); }
  const JacobianXiOplusType& jacobianOplusXi() const { return _jacobianOplusXi; }
 protected:
  JacobianXiOplusType _jacobianOplusXi;
};
}  // namespace g2o
