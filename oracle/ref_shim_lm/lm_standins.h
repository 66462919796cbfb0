// -*- C++ -*-
// TEST INFRASTRUCTURE (oracle/ref_shim_lm), force-included (-include) when oracle/Makefile compiles the reference's own
// Levenberg-Marquardt controller -- Thirdparty/g2o/g2o/core/optimization_algorithm_levenberg.cpp with
// optimization_algorithm_with_hessian.cpp, optimization_algorithm.cpp, solver.cpp, batch_stats.cpp and
// stuff/{property,string_tools,timeutil}.cpp, all UNMODIFIED -- into oracle/_ref/libg2o_ref_lm.so.
//
// Those sources need three g2o headers that in turn need the real Eigen and the whole graph machinery:
// sparse_block_matrix.h, optimizable_graph.h and sparse_optimizer.h.  The Makefile pre-defines their include guards and this
// header provides what the compiled sources touch instead: a declaration of SparseBlockMatrix (only named in signatures), and
// SparseOptimizer / OptimizableGraph as an abstract interface with exactly the members the controller calls
// (sparse_optimizer.h:87-303: computeActiveErrors, activeRobustChi2, push, pop, discardTop, update, terminate, indexMapping,
// activeVertices; optimizable_graph.h: Vertex::dimension, hessian(i, j), marginalized).  oracle/ref_lm_pin.cc implements that
// interface on the oracle's level-1 entry points, so the reference's controller drives the oracle's linear algebra.
#pragma once
#include <cstddef>
#include <iomanip>
#include <cstring>
#include <vector>
#include <Eigen/Core>
#include "Thirdparty/g2o/g2o/core/hyper_graph.h"

namespace g2o {
using namespace Eigen;
template <class MatrixType> class SparseBlockMatrix;

struct OptimizableGraph {
  class Vertex {
   public:
    virtual ~Vertex() {}
    virtual int dimension() const = 0;
    virtual const double& hessian(int i, int j) const = 0;
    virtual bool marginalized() const = 0;
  };
  typedef std::vector<Vertex*> VertexContainer;
};

class SparseOptimizer {
 public:
  virtual ~SparseOptimizer() {}
  virtual void computeActiveErrors() = 0;
  virtual double activeRobustChi2() const = 0;
  virtual void push() = 0;
  virtual void pop() = 0;
  virtual void discardTop() = 0;
  virtual void update(const double* update) = 0;
  virtual bool terminate() = 0;
  virtual const OptimizableGraph::VertexContainer& indexMapping() const = 0;
  virtual const OptimizableGraph::VertexContainer& activeVertices() const = 0;
};
}  // namespace g2o
