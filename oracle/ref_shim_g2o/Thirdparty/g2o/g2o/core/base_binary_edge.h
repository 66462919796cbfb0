// -*- C++ -*-
// TEST INFRASTRUCTURE (oracle/ref_shim): see base_vertex.h beside this file.
#pragma once
#include "base_vertex.h"
namespace g2o {
template <int D, class E, class VertexXi, class VertexXj> class BaseBinaryEdge : public BaseEdge<D, E> {
 public:
  typedef Eigen::Matrix<double, D, VertexXi::Dimension> JacobianXiOplusType;
  typedef Eigen::Matrix<double, D, VertexXj::Dimension> JacobianXjOplusType;
  BaseBinaryEdge() { this->_vertices.resize(2, nullptr); }
  const JacobianXiOplusType& jacobianOplusXi() const { return _jacobianOplusXi; }
  const JacobianXjOplusType& jacobianOplusXj() const { return _jacobianOplusXj; }
 protected:
  JacobianXiOplusType _jacobianOplusXi;
  JacobianXjOplusType _jacobianOplusXj;
};
}  // namespace g2o
