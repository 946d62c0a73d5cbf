// data.h -- compatibility layer for the reference's loader API (include/data.h, src/data.cu):
// load_graph / reorderCSR with the reference's signatures, implemented on gnnagg_graph_* and
// gnnagg_reorder_csr of libgnnagg.so.  File formats stay byte-compatible.
#ifndef DATA_H
#define DATA_H
#include "util.h"

template <class T>
T *createCudaMatrixCopy(T *d, int count)
{
    T *p_d = NULL;
    cudaMalloc2((void **)&p_d, count * sizeof(T));
    cudaMemcpy(p_d, d, count * sizeof(T), cudaMemcpyHostToDevice);
    return p_d;
}

// prints the first `outnum` values of a device buffer (data.h:21-38 of the reference)
template <class T>
void testGPUBuffer(int gpuid, T *dptr, int outnum = 64)
{
    checkCudaErrors(cudaSetDevice(gpuid));
    std::vector<T> host((size_t)outnum);
    checkCudaErrors(cudaMemcpy(host.data(), dptr, outnum * sizeof(T), cudaMemcpyDeviceToHost));
    printf("gpu %d\n", gpuid);
    for (int j = 0; j < outnum; ++j) {
        if (j % 32 == 0) cout << endl;
        cout << host[j] << ' ';
    }
    cout << '\n';
}

// the map[i]-th old vertex is placed at new position i (src/data.cu:4-29); allocates when NULL
inline void reorderCSR(const int *ptr, const int *idx, const int *map, const int *reverse_map, int num_v, int num_e,
                       int *&newptr, int *&newidx)
{
    if (newptr == NULL) newptr = new int[num_v + 1];
    if (newidx == NULL) newidx = new int[num_e > 0 ? num_e : 1];
    checkGnnagg(gnnagg_reorder_csr(ptr, idx, map, reverse_map, num_v, num_e, newptr, newidx));
}

// src/data.cu:31-139.  Reads ../data/<dset>.{config,graph[.ptrdump,.edgedump]} relative to the CWD;
// applies ../data/<dset>.reorder<subfix> when present.  Like the reference, an empty subfix falls
// back to the global `reorderfile` left by argParse (src/data.cu:96-98), and `rows` /
// `reverse_rows` are published as globals.
inline void load_graph(std::string dset, int &num_v, int &num_e, int *&indptr, int *&indices, bool shuffle = true,
                       std::string reorder_subfix = "")
{
    dbg("loading");
    const std::string basedir = "../data/";
    checkGnnagg(gnnagg_graph_config(basedir.c_str(), dset.c_str(), &num_v, &num_e));
    indptr = new int[num_v + 1];
    indices = new int[num_e > 0 ? num_e : 1];
    if (!reorder_subfix.empty()) reorderfile = basedir + dset + ".reorder" + reorder_subfix;
    const bool want = shuffle && reorderfile.size() > 1 && fexist(reorderfile);
    int *r = NULL, *rr = NULL, did = 0;
    if (want) {
        r = new int[num_v];
        rr = new int[num_v];
    }
    checkGnnagg(gnnagg_graph_load(basedir.c_str(), dset.c_str(), want ? reorderfile.c_str() : NULL, num_v, num_e,
                                  indptr, indices, r, rr, &did));
    if (did) {
        dbg("reorder:" + reorderfile);
        rows = r;
        reverse_rows = rr;
    } else {
        dbg("unreordered");
        dbg(reorderfile);
    }
}
#endif
