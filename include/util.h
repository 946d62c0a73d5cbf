// util.h -- compatibility layer: the support types and globals the reference's aggregator API is
// expressed in (reference include/util.h + src/util.cu), re-implemented header-only on top of
// libgnnagg.so so that the reference's drivers (Figure9/main.cu, Figure10/main_a.cu, main_b.cu)
// compile unmodified against this include directory.  Needs -std=c++17 (inline variables).
//
// Kept: names, signatures and observable behaviour -- flag names of argParse (src/util.cu:24-147),
// 512-byte rounding of cudaMalloc2 (util.h:144-152), registerPtr/safeFree ownership protocol
// (util.h:154-177), abort-on-error of checkCudaErrors (util.h:82-104), CSRSubGraph.
// Not kept: the vendored args.hxx parser (a 60-line flag loop replaces it) and the multi-GPU
// leftovers that no driver touches.
#ifndef UTIL_H
#define UTIL_H

#include <assert.h>
#include <cublas_v2.h>
#include <cuda_runtime.h>
#include <curand.h>
#include <sys/stat.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <sstream>
#include <string>
#include <algorithm>
#include <vector>

#include "dbg.h"
#include "gnnagg.h"

using namespace std;  // the reference headers do this and its drivers rely on it (util.h:29)

#define CEIL(a, b) (((a) + (b)-1) / (b))

// ---- mutable global state of the reference (src/util.cu:3-22,153-178), header-only here ----
inline int GPUNUM = 1;
inline int NEINUM = -1;
inline int n = -1, m = -1, feature_len = 0;
inline int *rows = NULL;
inline int *reverse_rows = NULL;
inline vector<void *> registered_ptr;
inline int outfea = 0;
inline int total_size = 0;
inline string inputfeature, inputweight, inputgraph, edgefile, ptrfile, partitionfile, reorderfile, inputtransgraph,
    partialgraphs;
inline int *gptr = NULL, *gidx = NULL;
inline float *gval = NULL;
inline cublasHandle_t cublasH = NULL;
inline cudaStream_t stream = NULL;
inline int *numVertex = new int[1], *numEdge = new int[1];
inline int **gptrs = NULL, **gidxs = NULL;
inline cublasHandle_t *cublasHs = new cublasHandle_t[1]();

inline bool fexist(const std::string &name)
{
    struct stat st;
    return stat(name.c_str(), &st) == 0;
}

#define timestamp(__var__) auto __var__ = std::chrono::system_clock::now();

inline double getDuration(std::chrono::time_point<std::chrono::system_clock> a,
                          std::chrono::time_point<std::chrono::system_clock> b)
{
    return std::chrono::duration<double>(b - a).count();
}

// 2*m*F/t/1e9, in double (the reference multiplies in int and overflows on reddit, util.h:125)
inline double getFLOP(double time)
{
    assert(time > 0 && m > 0 && feature_len > 0);
    return 2.0 * (double)m * (double)feature_len / time / 1e9;
}

namespace gnnagg_compat {
// prints the message with its source position, resets the device and terminates: the reference's only error
// policy (util.h:82-104 there), kept because its drivers assume a failed call never returns
[[noreturn]] inline void die(const std::string &what, const char *file, int line)
{
    std::cerr << what << "\n" << file << ':' << line << "\nAborting...\n";
    cudaDeviceReset();
    exit(1);
}
inline void expect_cuda(long status, const char *file, int line)
{
    if (status != 0) die("Cuda failure: " + std::to_string(status), file, line);
}
inline void expect_gnnagg(int status, const char *file, int line)
{
    if (status != 0) die(std::string("gnnagg failure: ") + gnnagg_last_error(), file, line);
}
}  // namespace gnnagg_compat

#define FatalError(s) gnnagg_compat::die(std::string(s), __FILE__, __LINE__)
#define checkCudaErrors(status) gnnagg_compat::expect_cuda((long)(status), __FILE__, __LINE__)
#define checkGnnagg(status) gnnagg_compat::expect_gnnagg((status), __FILE__, __LINE__)  // same policy for the C ABI

static inline unsigned int roundUp(unsigned int nominator, unsigned int denominator)
{
    return (nominator + denominator - 1) / denominator;
}

// ---- allocation helpers of the reference API (util.h:144-195 there): same names and behaviour, bodies written
// against two small primitives of this shim (rounded allocation, upload of a host range)
namespace gnnagg_compat {
constexpr size_t kAllocQuantum = 512;  // the reference rounds every device allocation up to 512 bytes

inline cudaError_t alloc_rounded(void **out, size_t bytes)
{
    const size_t quanta = (bytes + kAllocQuantum - 1) / kAllocQuantum;
    return cudaMalloc(out, quanta * kAllocQuantum);
}

// device copy of host[0, count): one rounded allocation + one blocking H2D copy, aborting like the reference on failure
template <class T>
inline T *upload(const T *host, size_t count)
{
    void *dev = NULL;
    const size_t bytes = count * sizeof(T);
    if (bytes != 0) {
        total_size += (int)bytes;
        expect_cuda((long)alloc_rounded(&dev, bytes), __FILE__, __LINE__);
        expect_cuda((long)cudaMemcpy(dev, host, bytes, cudaMemcpyHostToDevice), __FILE__, __LINE__);
    }
    return static_cast<T *>(dev);
}

inline bool externally_owned(const void *p)
{
    return std::find(registered_ptr.begin(), registered_ptr.end(), const_cast<void *>(p)) != registered_ptr.end();
}
}  // namespace gnnagg_compat

inline cudaError_t cudaMalloc2(void **a, size_t s)
{
    if (s == 0) return cudaSuccess;  // pointer left untouched, as in the reference
    total_size += (int)s;
    return gnnagg_compat::alloc_rounded(a, s);
}

template <class T>
inline void registerPtr(T ptr)
{
    registered_ptr.push_back(reinterpret_cast<void *>(ptr));
}

// releases a device pointer the aggregator owns; pointers announced through registerPtr stay with their owner
template <class T>
void safeFree(T *&a)
{
    if (a == NULL || gnnagg_compat::externally_owned(a)) return;
    cudaFree(a);
    cudaGetLastError();  // a stale pointer must not poison later launches
    a = NULL;
}

template <class T>
T *createCopy(T *p, int size)
{
    return gnnagg_compat::upload(p, (size_t)size);
}

template <class T>
void copyVec2Dev(std::vector<T> *vec, T *&output)
{
    assert(output == NULL);
    output = gnnagg_compat::upload(vec->data(), vec->size());
    std::vector<T>().swap(*vec);  // the host vector gives its storage back
}

struct CSR {
    CSR(int *outptr, int *outidx, float *outval) : ptr(outptr), idx(outidx), val(outval) {}
    int *ptr;
    int *idx;
    float *val;
};

class CSRSubGraph {
public:
    CSRSubGraph(int *outvertexset, int *outptr, int *outidx, int vertex_num, int edge_num)
        : vertexset(outvertexset), ptr(outptr), idx(outidx), num_v(vertex_num), num_e(edge_num)
    {
    }
    void free()
    {
        safeFree(vertexset);
        safeFree(ptr);
        safeFree(idx);
    }
    int *vertexset = NULL;
    int *ptr = NULL;
    int *idx = NULL;
    int num_v = 0;
    int num_e = 0;
};

// Command line of the reference drivers (src/util.cu:24-147): --dataset (required), --datadir,
// --partition-path, --reorder <suffix>, --gpu-num, --nei, --feature-len (required), --outfea,
// --limit, --limit2; both "--flag value" and "--flag=value".  Fills the globals above.
inline void argParse(int argc, char **argv, int *p_limit = NULL, int *p_limit2 = NULL)
{
    string dset, ddir = "../data/", reorder_suffix;
    bool has_dset = false, has_feat = false, has_reorder = false, has_limit = false, has_limit2 = false;
    for (int i = 1; i < argc; ++i) {
        string a = argv[i], value;
        if (a.rfind("--", 0) != 0) {
            std::cerr << "unexpected argument " << a << std::endl;
            exit(1);
        }
        const size_t eq = a.find('=');
        if (eq != string::npos) {
            value = a.substr(eq + 1);
            a = a.substr(0, eq);
        } else if (i + 1 < argc) {
            value = argv[++i];
        } else {
            std::cerr << "flag " << a << " needs a value" << std::endl;
            exit(1);
        }
        const string key = a.substr(2);
        if (key == "dataset") dset = value, has_dset = true;
        else if (key == "datadir") ddir = value;
        else if (key == "partition-path") partitionfile = value;
        else if (key == "reorder") reorder_suffix = value, has_reorder = true;
        else if (key == "gpu-num") GPUNUM = atoi(value.c_str());
        else if (key == "nei") NEINUM = atoi(value.c_str());
        else if (key == "feature-len") feature_len = atoi(value.c_str()), has_feat = true;
        else if (key == "outfea") outfea = atoi(value.c_str());
        else if (key == "limit") { if (p_limit) *p_limit = atoi(value.c_str()); has_limit = true; }
        else if (key == "limit2") { if (p_limit2) *p_limit2 = atoi(value.c_str()); has_limit2 = true; }
        else {
            std::cerr << "unknown flag " << a << std::endl;
            exit(1);
        }
    }
    assert(has_dset);
    assert(has_feat);
    if (!partitionfile.empty()) assert(fexist(partitionfile));
    if (int rc = gnnagg_graph_config(ddir.c_str(), dset.c_str(), &n, &m)) {
        (void)rc;
        FatalError(string("config: ") + gnnagg_last_error());
    }
    inputgraph = ddir + dset + ".graph";
    ptrfile = inputgraph + ".ptrdump";
    edgefile = inputgraph + ".edgedump";
    if (fexist(inputgraph)) {
        if (!fexist(ptrfile)) ptrfile = "";
        if (!fexist(edgefile)) edgefile = "";
    } else {
        assert(fexist(ptrfile) && fexist(edgefile));
    }
    reorderfile = ddir + dset + ".reorder";
    if (reorder_suffix.size() > 1) reorderfile += reorder_suffix;
    if (has_reorder)
        assert(fexist(reorderfile));
    else
        reorderfile = "";
    if (p_limit) assert(has_limit);
    if (p_limit2) assert(has_limit2);
    dbg(dset);
    inputgraph = dset;  // the drivers pass the bare dataset name on to load_graph (src/util.cu:133)
}

#endif
