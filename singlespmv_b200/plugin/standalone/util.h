// util.h -- standalone stand-in for the reference's src/util.h, used when the plugin is built outside
// the reference tree.  Same type names, same field order (src/util.h:7-28), same function names and
// meaning (src/util.cpp); written from the interface, not copied.
#pragma once
#include <string>

struct SpMat {                 // COO sorted by (row, col)
    int nRow, nCol, nNnz;
    int *row_idx;
    int *col_idx;
    double *val;
};
struct Vec {
    int size;
    double *val;
};
void LoadSparseMatrix (SpMat &A, const std::string &matFile);      // src/util.cpp:30-66 semantics
double GetTimeBySec ();                                            // src/util.cpp:21-25
bool VerifyResult (const SpMat &A, const Vec &x, const Vec &y);    // src/util.cpp:67-83: abs <= 1e-6 OR rel <= 1e-6
std::string GetBasename (const std::string &path);
Vec CreateRandomVector (int size);                                 // src/util.cpp:92-102: rand()/RAND_MAX
