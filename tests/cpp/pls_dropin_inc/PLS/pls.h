// Stand-in for the reference's <PLS/pls.h> when a program written against the PLS library is built on the CUDA library instead
// (INTEGRATION.md §3): the public typedefs of lib/PLS/include/PLS/pls.h:21-33 (Eigen comes from oracle/shim/ in this image; real
// Eigen works the same way) followed by abc_b200.hpp's ABCB200_DROP_IN_PLS block, which defines namespace PLS with the reference's
// names. With this directory in front on the include path the reference's lib/PLS/src/main.cpp compiles UNMODIFIED
// (tests/cpp/Makefile: _build/pls_main_dropin).
#ifndef ABCB200_PLS_DROPIN_PLS_H
#define ABCB200_PLS_DROPIN_PLS_H
#include <Eigen/Core>
#include <vector>
#include <iostream>
#include <random>
#include <algorithm>
#include <numeric>
typedef double float_type;
typedef Eigen::MatrixXd Mat2D;
typedef Eigen::VectorXd Col;
typedef Eigen::RowVectorXd Row;
typedef Eigen::VectorXi Coli;
typedef Eigen::Matrix<size_t, Eigen::Dynamic, 1> Colsz;
typedef Eigen::RowVectorXi Rowi;
typedef Eigen::Matrix<size_t, 1, Eigen::Dynamic> Rowsz;
#define ABCB200_DROP_IN_PLS
#include "../../../../abcsmc_b200/host/abc_b200.hpp"
#endif
