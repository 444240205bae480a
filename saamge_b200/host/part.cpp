// Graph partitioning input producer: METIS k-way + connected-components pass.
// Mirrors part_generate_partitioning / connectedComponents
// (amg/src/part.cpp:56-118, 120-215).  METIS is the static library shipped with
// the CUDA toolkit (libmetis_static.a, 64-bit idx_t, no header), so prototypes
// and option indices (METIS 5.1 layout) are declared here by hand.
#include <mutex>
#include "part.hpp"

#include <algorithm>
#include <cstring>

extern "C" {
int METIS_SetDefaultOptions(int64_t *options);
int METIS_PartGraphKway(int64_t *nvtxs, int64_t *ncon, int64_t *xadj,
                        int64_t *adjncy, int64_t *vwgt, int64_t *vsize,
                        int64_t *adjwgt, int64_t *nparts, float *tpwgts,
                        float *ubvec, int64_t *options, int64_t *objval,
                        int64_t *part);
}

namespace saamge
{

enum
{
    SA_METIS_OPTION_PTYPE = 0,
    SA_METIS_OPTION_CONTIG = 11,
    SA_METIS_OPTION_UFACTOR = 16,
    SA_METIS_OPTION_NUMBERING = 17,
    SA_METIS_NOPTIONS = 40,
    SA_METIS_PTYPE_KWAY = 1,
    SA_METIS_OK = 1
};

int connectedComponents(std::vector<int> &partitioning, const Table &conn)
{
    const int num_nodes = conn.Size();
    int num_part = 0;
    for (int i = 0; i < num_nodes; ++i)
        num_part = std::max(num_part, partitioning[i] + 1);

    std::vector<int> component(num_nodes, -1);
    std::vector<int> offset_comp((size_t)num_part + 1, 0);
    int *num_comp = offset_comp.data() + 1;
    const int *i_table = conn.GetI();
    const int *j_table = conn.GetJ();
    std::vector<int> vertex_stack(num_nodes);
    int stack_p = 0, stack_top_p = 0;
    for (int node = 0; node < num_nodes; node++)
    {
        if (partitioning[node] < 0)
            continue;
        if (component[node] >= 0)
            continue;
        component[node] = num_comp[partitioning[node]]++;
        vertex_stack[stack_top_p++] = node;
        for (; stack_p < stack_top_p; stack_p++)
        {
            const int i = vertex_stack[stack_p];
            if (partitioning[i] < 0)
                continue;
            for (int j = i_table[i]; j < i_table[i + 1]; j++)
            {
                const int k = j_table[j];
                if (partitioning[k] == partitioning[i] && component[k] < 0)
                {
                    component[k] = component[i];
                    vertex_stack[stack_top_p++] = k;
                }
            }
        }
    }
    for (int p = 0; p < num_part; ++p)
        offset_comp[p + 1] += offset_comp[p];
    for (int i = 0; i < num_nodes; ++i)
        partitioning[i] = offset_comp[partitioning[i]] + component[i];
    return offset_comp[num_part];
}

int *part_generate_partitioning(const Table &graph, const int *weights, int *parts)
{
    SA_ASSERT(graph.Size() == graph.Width() || *parts == 1);
    const int nodes_number = graph.Size();
    int *partitioning = new int[nodes_number];
    const int target_parts = *parts;
    SA_ASSERT(target_parts > 0);

    if (target_parts > 1)
    {
        int64_t options[SA_METIS_NOPTIONS];
        METIS_SetDefaultOptions(options);
        options[SA_METIS_OPTION_PTYPE] = SA_METIS_PTYPE_KWAY;
        options[SA_METIS_OPTION_NUMBERING] = 0;
        options[SA_METIS_OPTION_CONTIG] = 1;
        options[SA_METIS_OPTION_UFACTOR] = 30;
        int64_t nvtxs = nodes_number, ncon = 1, nparts = target_parts, objval = 0;
        std::vector<int64_t> xadj(graph.I.begin(), graph.I.end());
        std::vector<int64_t> adjncy(graph.J.begin(), graph.J.end());
        std::vector<int64_t> vwgt((size_t)nodes_number);
        for (int i = 0; i < nodes_number; ++i)
            vwgt[i] = weights ? weights[i] : 1;
        std::vector<int64_t> part64((size_t)nodes_number);
        // METIS calls are serialised: the coarse topology of the next level is prefetched on a
        // helper thread (ml.cpp) and the thread safety of this METIS build is unknown
        static std::mutex metis_mutex;
        std::lock_guard<std::mutex> metis_lock(metis_mutex);
        const int stat = METIS_PartGraphKway(&nvtxs, &ncon, xadj.data(), adjncy.data(),
                                             vwgt.data(), NULL, NULL, &nparts, NULL, NULL,
                                             options, &objval, part64.data());
        SA_ASSERT(SA_METIS_OK == stat);
        for (int i = 0; i < nodes_number; ++i)
            partitioning[i] = (int)part64[i];
    }
    else
        std::memset(partitioning, 0, sizeof(*partitioning) * nodes_number);

    if (target_parts > 1)
    {
        std::vector<int> p_array(partitioning, partitioning + nodes_number);
        connectedComponents(p_array, graph);
        std::copy(p_array.begin(), p_array.end(), partitioning);
    }
    int actual_parts = 0;
    for (int i = 0; i < nodes_number; ++i)
        actual_parts = std::max(actual_parts, partitioning[i] + 1);
    *parts = actual_parts;
    return partitioning;
}

int *part_generate_partitioning_unweighted(const Table &graph, int *parts)
{
    return part_generate_partitioning(graph, NULL, parts);
}

int *part_generate_partitioning_blocks(int dim, int nx, int ny, int nz, int bx,
                                       int by, int bz, int *parts)
{
    if (dim == 2)
    {
        nz = 1;
        bz = 1;
    }
    const int px = (nx + bx - 1) / bx, py = (ny + by - 1) / by,
              pz = (nz + bz - 1) / bz;
    int *partitioning = new int[(size_t)nx * ny * nz];
    for (int k = 0; k < nz; ++k)
        for (int j = 0; j < ny; ++j)
            for (int i = 0; i < nx; ++i)
                partitioning[(size_t)i + (size_t)nx * (j + (size_t)ny * k)] =
                    (i / bx) + px * ((j / by) + py * (k / bz));
    *parts = px * py * pz;
    return partitioning;
}

} // namespace saamge
