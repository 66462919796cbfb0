// -*- C++ -*-
// TEST INFRASTRUCTURE (oracle/ref_shim): see base_vertex.h beside this file.  g2o sizes the Jacobian of vertex i as
// D x dimension(vertex i) when it maps the workspace (base_multi_edge.hpp, mapHessianMemory/linearizeOplus); here the
// block is sized when the vertex is attached.
#pragma once
#include "base_vertex.h"
namespace g2o {
template <int D, class E> class BaseMultiEdge : public BaseEdge<D, E> {
 public:
  typedef Eigen::MatrixXd JacobianType;
  void resize(size_t n) { this->_vertices.resize(n, nullptr); _jacobianOplus.resize(n); }
  void setVertex(size_t i, OptimizableGraph::Vertex* v) override {
    this->_vertices[i] = v;
    _jacobianOplus[i].resize(D, v->dimension());
  }
  const std::vector<JacobianType>& jacobianOplus() const { return _jacobianOplus; }
 protected:
  std::vector<JacobianType> _jacobianOplus;
};
}  // namespace g2o
