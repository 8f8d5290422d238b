// main_b200.cpp -- the reference's driver loop (src/main.cpp:17-209) around the B200 plugin:
// load -> random x, y (srand(3), x first) -> OptimizeProblem -> 2x {SpMV; VerifyResult} -> calibrate
// `loop` by doubling until >= 1 s -> 10 tries of `loop` calls, keep the minimum -> key/value report between
// the `++++` / `----` lines (parseable by the reference's log/format.cpp:32-49).
//
//     spmv_b200 <matrix.mtx>                       the reference's own invocation
//     spmv_b200 synth:<kind>:<p0>[:<p1>[:<seed>]]  kind = lap2d5 | lap3d7 | box3d27 | uniform | rmat
//                                                  (BASELINE.json shapes are generated in HBM; a 938 M-entry
//                                                   Matrix-Market text file is not practical)
// Built once per format, like the reference builds one binary per -DOPT_<FMT> (Makefile:10-21).
// With -DB200_DEVICE_RESIDENT each try is timed with CUDA events around the `loop` back-to-back launches;
// otherwise with the reference's wall clock (the plugin then synchronises inside every SpMV).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>
#include <cuda_runtime_api.h>
#include "opt_b200.h"
#include "util.h"

extern std::vector<int> g_step_count;                                 // filled by the SS plugin (src/opt_ss.cpp:143-147)
extern std::vector<double> g_step_time;
extern std::vector<double> g_profile;

#define DRV_CUDA(e) do { cudaError_t e_ = (e); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #e, cudaGetErrorString(e_)); exit(1); } } while (0)

static bool parse_synth (const std::string &arg, SpMat &A) {
    if (arg.compare(0, 6, "synth:") != 0) return false;
    std::vector<std::string> tok;
    size_t at = 6;
    while (at <= arg.size()) {
        size_t e = arg.find(':', at);
        if (e == std::string::npos) e = arg.size();
        tok.push_back(arg.substr(at, e - at));
        at = e + 1;
    }
    static const char *kinds[] = {"lap2d5", "lap3d7", "box3d27", "uniform", "rmat"};
    int kind = -1;
    for (int k = 0; k < 5; k++) if (!tok.empty() && tok[0] == kinds[k]) kind = k;
    if (kind < 0 || tok.size() < 2) { fprintf(stderr, "bad synth spec '%s'\n", arg.c_str()); exit(1); }
    const long long p0 = atoll(tok[1].c_str()), p1 = tok.size() > 2 ? atoll(tok[2].c_str()) : 0;
    const unsigned long long seed = tok.size() > 3 ? strtoull(tok[3].c_str(), NULL, 10) : (kind == 4 ? 42 : 1);
    b200spmv_coo coo;
    if (b200spmv_synth(kind, p0, p1, seed, 0, 0, &coo, NULL) != 0) { fprintf(stderr, "synth: %s\n", b200spmv_last_error()); exit(1); }
    A.nRow = coo.nRow; A.nCol = coo.nCol; A.nNnz = (int)coo.nnz;
    A.row_idx = new int[coo.nnz > 0 ? coo.nnz : 1];
    A.col_idx = new int[coo.nnz > 0 ? coo.nnz : 1];
    A.val = new double[coo.nnz > 0 ? coo.nnz : 1];
    if (b200spmv_coo_download(&coo, A.row_idx, A.col_idx, A.val) != 0) { fprintf(stderr, "download: %s\n", b200spmv_last_error()); exit(1); }
    b200spmv_coo_free(&coo);
    return true;
}

int main (int argc, char **argv) {
    srand(3);                                                             // src/main.cpp:18
    if (argc < 2) {
        printf("Usage: %s <matrix.mtx | synth:kind:p0[:p1[:seed]]>\n", argv[0]);
        exit(1);
    }
    const std::string matFile = argv[1];
    SpMat A;
    std::cerr << "Loading sparse matrix " << matFile << " ... ";
    if (!parse_synth(matFile, A)) LoadSparseMatrix(A, matFile);
    std::cerr << "done." << std::endl;
    const int nRow = A.nRow, nCol = A.nCol, nNnz = A.nNnz;
    Vec x = CreateRandomVector(nCol);                                     // x before y: src/main.cpp:31-32
    Vec y = CreateRandomVector(nRow);
    SpMatOpt A_opt;
    VecOpt x_opt;
    std::cerr << "Optimizing ... ";
    const double tConv = -GetTimeBySec();
    OptimizeProblem(A, x, A_opt, x_opt);
    const double convertMs = (tConv + GetTimeBySec()) * 1e3;
    std::cerr << "done." << std::endl;

#ifndef B200_NO_VERIFY
    for (int i = 0; i < 2; i++) {                                         // src/main.cpp:40-56
        g_profile = std::vector<double>(10);
        SpMV(A_opt, x_opt, y);
        B200FetchResult(A_opt, y);
        std::cerr << "Verifying " << i << " ... ";
        if (!VerifyResult(A, x, y)) {
            printf("*** invalid result ***\n");
            exit(1);
        }
        std::cerr << "done." << std::endl;
    }
#endif

    // the reference hard-codes 1 s of calibration and 10 tries; the environment can shorten both (tests)
    const double minSeconds = getenv("SPMV_MIN_SECONDS") ? atof(getenv("SPMV_MIN_SECONDS")) : 1.0;
    const int nTry = getenv("SPMV_NTRY") ? atoi(getenv("SPMV_NTRY")) : 10;
    int loop = 1;
    std::cerr << "Calculating SpMV ... ";
    {
        g_profile = std::vector<double>(10);
        const double t0 = GetTimeBySec();                                 // src/main.cpp:58-71
        do {
            for (int i = 0; i < loop; i++) SpMV(A_opt, x_opt, y);
            B200Synchronize(A_opt);
            loop *= 2;
        } while (GetTimeBySec() - t0 < minSeconds);
    }
    double minElapsedTime = 0;
    std::vector<double> g_best_profile;
    {
        cudaEvent_t e0, e1;
        DRV_CUDA(cudaEventCreate(&e0));
        DRV_CUDA(cudaEventCreate(&e1));
        // src/main.cpp:79-102
        for (int t = 0; t < nTry; t++) {
            g_profile = std::vector<double>(10);
            double elapsed;
#ifdef B200_DEVICE_RESIDENT
            if (A_opt.mg) {                                               // several GPUs: wall clock around the asynchronous loop
                B200Synchronize(A_opt);
                elapsed = -GetTimeBySec();
                for (int i = 0; i < loop; i++) SpMV(A_opt, x_opt, y);
                B200Synchronize(A_opt);
                elapsed += GetTimeBySec();
                elapsed /= loop;
                if (t == 0 || elapsed < minElapsedTime) { minElapsedTime = elapsed; g_best_profile = g_profile; }
                continue;
            }
            DRV_CUDA(cudaEventRecord(e0, (cudaStream_t)A_opt.stream));
            for (int i = 0; i < loop; i++) SpMV(A_opt, x_opt, y);
            DRV_CUDA(cudaEventRecord(e1, (cudaStream_t)A_opt.stream));
            DRV_CUDA(cudaEventSynchronize(e1));
            float ms = 0;
            DRV_CUDA(cudaEventElapsedTime(&ms, e0, e1));
            elapsed = ms * 1e-3 / loop;
#else
            elapsed = -GetTimeBySec();
            for (int i = 0; i < loop; i++) SpMV(A_opt, x_opt, y);
            elapsed += GetTimeBySec();
            elapsed /= loop;
#endif
            if (t == 0 || elapsed < minElapsedTime) {
                minElapsedTime = elapsed;
                g_best_profile = g_profile;
            }
        }
    }
    std::cerr << "done." << std::endl;

    int dev = 0;
    cudaDeviceProp prop;
    DRV_CUDA(cudaGetDevice(&dev));
    DRV_CUDA(cudaGetDeviceProperties(&prop, dev));
    const long long algBytes = B200Scalar(A_opt, "alg_bytes");
    const double gbs = algBytes / minElapsedTime / 1e9;
    printf("++++++++++++++++++++++++++++++++++++++++\n");                 // src/main.cpp:108-207
    printf("%25s\t%s\n", "Architecture", "GPU");
    printf("%25s\t%s\n", "MatrixFormat", B200FormatName());
    printf("%25s\t%s\n", "Device", prop.name);
    if (!strcmp(B200FormatName(), "SS")) {                                // src/main.cpp:155-162
        printf("%25s\t%d\n", "nStep", int(g_step_count.size()));
        for (size_t i = 0; i < g_step_count.size(); i++) printf("%22s-%02d\t%d\n", "StepCount", int(i), g_step_count[i]);
    }
#if defined(SEGMENT_WIDTH)
    printf("%25s\t%d\n", "SEGMENT_WIDTH(byte)", int(SEGMENT_WIDTH * sizeof(double)));
#endif
#if defined(PROFILING) && defined(B200_SS_FAITHFUL)
    if (g_best_profile.size() >= 2 && g_best_profile[0] > 0 && g_best_profile[1] > 0) {   // src/main.cpp:171-174,184-187
        const bool css = !strcmp(B200FormatName(), "CSS");
        printf("%25s\t%lf\n", css ? "MulPerf(GFLOPS)" : "MulPerf", 2.0 * nNnz / (g_best_profile[0] / loop) / 1e9);
        printf("%25s\t%lf\n", css ? "SumPerf(GFLOPS)" : "SumPerf", 2.0 * nNnz / (g_best_profile[1] / loop) / 1e9);
    }
#endif
#if defined(N_BLOCK)
    printf("%25s\t%d\n", "N_BLOCK", N_BLOCK);
#endif
    printf("%25s\t%s\n", "Matrix", GetBasename(matFile).c_str());
    printf("%25s\t%s\n", "MatrixPath", matFile.c_str());
    printf("%25s\t%lf\n", "Performance(GFLOPS)", 2.0 * nNnz / minElapsedTime / 1e9);   // 64-bit, src/main.cpp:196 overflows
    printf("%25s\t%d\n", "nRow", nRow);
    printf("%25s\t%d\n", "nCol", nCol);
    printf("%25s\t%d\n", "nNnz", nNnz);
    printf("%25s\t%d\n", "nThread", 1);                                   // host threads driving the device (src/main.cpp:200-206)
    printf("%25s\t%d\n", "nGPU", A_opt.nGPU);
    if (A_opt.mg) {
        printf("%25s\t%lld\n", "HaloDoublesPerSpMV", B200Scalar(A_opt, "halo_total"));
        printf("%25s\t%lld\n", "GraphLaunch", B200Scalar(A_opt, "graphed"));
    }
    printf("%25s\t%lf\n", "KernelTime(us)", minElapsedTime * 1e6);
    printf("%25s\t%lld\n", "AlgBytes", algBytes);
    printf("%25s\t%lf\n", "EffectiveBW(GB/s)", gbs);
    printf("%25s\t%lf\n", "RooflinePct(8000GB/s)", 100.0 * gbs / (8000.0 * (A_opt.nGPU > 0 ? A_opt.nGPU : 1)));   // per GPU
    printf("%25s\t%lld\n", "LaunchesPerSpMV", B200Scalar(A_opt, "launches"));
    printf("%25s\t%lf\n", "ConvertTime(ms)", convertMs);
#ifdef B200_DEVICE_RESIDENT
    printf("%25s\t%s\n", "VectorResidency", "device");
#else
    printf("%25s\t%s\n", "VectorResidency", "host(H2D+D2H per SpMV)");
#endif
    printf("----------------------------------------\n");
    return 0;
}
