// util.cpp -- the driver-side helpers of plugin/standalone/util.h, restated from the behaviour of the reference's
// src/util.cpp (loader: first line not starting with '%' is "M N L", then exactly L "row col val"
// triples, 1-based, sorted by (row, col), duplicates kept, banner/symmetry ignored).
#include "util.h"
#include <vector>

// report globals written by the SS / CSS plugins (src/util.cpp:16-18)
std::vector<int> g_step_count;
std::vector<double> g_step_time;
std::vector<double> g_profile;

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <numeric>
#include <sstream>
#include <vector>
#include <sys/time.h>

double GetTimeBySec () {
    timeval tv;
    gettimeofday(&tv, NULL);
    return tv.tv_sec + tv.tv_usec * 1e-6;
}
std::string GetBasename (const std::string &path) {
    return path.substr(path.rfind('/') + 1);
}
void LoadSparseMatrix (SpMat &A, const std::string &matFile) {
    std::ifstream in(matFile.c_str());
    if (!in.is_open()) {
        std::cerr << "File not Found" << std::endl;
        exit(1);
    }
    std::string line;
    do std::getline(in, line); while (!line.empty() && line[0] == '%');
    int M = 0, N = 0, L = 0;
    std::stringstream(line) >> M >> N >> L;
    std::vector<int> r(L), c(L), order(L);
    std::vector<double> v(L);
    for (int i = 0; i < L; i++) {
        in >> r[i] >> c[i] >> v[i];
        r[i]--; c[i]--;
    }
    std::iota(order.begin(), order.end(), 0);
    std::sort(order.begin(), order.end(), [&](int a, int b) { return r[a] != r[b] ? r[a] < r[b] : c[a] < c[b]; });
    A.nRow = M; A.nCol = N; A.nNnz = L;
    A.row_idx = new int[L > 0 ? L : 1];
    A.col_idx = new int[L > 0 ? L : 1];
    A.val = new double[L > 0 ? L : 1];
    for (int i = 0; i < L; i++) {
        A.row_idx[i] = r[order[i]];
        A.col_idx[i] = c[order[i]];
        A.val[i] = v[order[i]];
    }
}
bool VerifyResult (const SpMat &A, const Vec &x, const Vec &y) {
    std::vector<double> res(A.nRow > 0 ? A.nRow : 1, 0.0);
    for (int i = 0; i < A.nNnz; i++) res[A.row_idx[i]] += A.val[i] * x.val[A.col_idx[i]];
    for (int i = 0; i < A.nRow; i++) {
        const double abs_err = fabs(res[i] - y.val[i]), rel_err = fabs(abs_err / res[i]);
        if (abs_err > 1e-6 && rel_err > 1e-6) {
            fprintf(stderr, "Error: %lf != %lf\n", res[i], y.val[i]);
            return false;
        }
    }
    return true;
}
Vec CreateRandomVector (int size) {
    Vec x;
    x.size = size;
    x.val = (double *)aligned_alloc(64, ((sizeof(double) * (size > 0 ? size : 1) + 63) / 64) * 64);
    for (int i = 0; i < size; i++) x.val[i] = double(rand()) / RAND_MAX;
    return x;
}
