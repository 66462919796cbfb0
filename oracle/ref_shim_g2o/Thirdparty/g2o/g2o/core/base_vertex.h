// -*- C++ -*-
// TEST INFRASTRUCTURE (oracle/ref_shim): the few members of g2o's vertex/edge base classes that the reference's
// include/G2oTypes.h and src/G2oTypes.cc touch (g2o/core/base_vertex.h, base_edge.h, base_unary_edge.h,
// base_binary_edge.h, base_multi_edge.h), as plain containers: no graph, no workspace maps, no robust kernel.  The vendored
// g2o needs the real Eigen.  With these the reference's edge classes compile unmodified and their computeError() /
// linearizeOplus() run on values set by oracle/ref_pin.cc.
#pragma once
#include <iostream>
#include <vector>
#include <Eigen/Core>

namespace g2o {

struct OptimizableGraph {
  struct Vertex {
    virtual ~Vertex() {}
    virtual int dimension() const = 0;
    virtual void oplusImpl(const double* update) = 0;
    virtual void setToOriginImpl() = 0;
    virtual bool read(std::istream& is) = 0;
    virtual bool write(std::ostream& os) const = 0;
    void oplus(const double* update) { oplusImpl(update); }
    void updateCache() {}
  };
};

template <int D, class T> class BaseVertex : public OptimizableGraph::Vertex {
 public:
  typedef T EstimateType;
  static const int Dimension = D;
  int dimension() const override { return D; }
  const EstimateType& estimate() const { return _estimate; }
  void setEstimate(const EstimateType& et) { _estimate = et; updateCache(); }
 protected:
  EstimateType _estimate;
};

template <int D, class E> class BaseEdge {
 public:
  static const int Dimension = D;
  typedef E Measurement;
  typedef Eigen::Matrix<double, D, 1> ErrorVector;
  typedef Eigen::Matrix<double, D, D> InformationType;
  virtual ~BaseEdge() {}
  virtual void computeError() = 0;
  virtual void linearizeOplus() = 0;
  virtual bool read(std::istream& is) = 0;
  virtual bool write(std::ostream& os) const = 0;
  void setMeasurement(const Measurement& m) { _measurement = m; }
  const Measurement& measurement() const { return _measurement; }
  void setInformation(const InformationType& i) { _information = i; }
  const InformationType& information() const { return _information; }
  const ErrorVector& error() const { return _error; }
  virtual void setVertex(size_t i, OptimizableGraph::Vertex* v) { _vertices[i] = v; }
  const std::vector<OptimizableGraph::Vertex*>& vertices() const { return _vertices; }
 protected:
  Measurement _measurement;
  InformationType _information;
  ErrorVector _error;
  std::vector<OptimizableGraph::Vertex*> _vertices;
};

}  // namespace g2o
