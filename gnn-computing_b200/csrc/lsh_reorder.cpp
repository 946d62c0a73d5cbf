// lsh_reorder.cpp -- placeholder until the clustering pass lands (see gnnagg.h: gnnagg_lsh_reorder)
#include "gnnagg.h"
#include "internal.h"
extern "C" int gnnagg_lsh_reorder(const int *, const int *, int, int, int, int, int, int, uint64_t, int *)
{
    return gnnagg::set_error(GNNAGG_ERR_STATE, "gnnagg_lsh_reorder: not built yet");
}
