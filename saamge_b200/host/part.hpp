// METIS wrapper + connected components (amg/inc/part.hpp, amg/src/part.cpp:56-215).
#ifndef SAAMGE_B200_PART_HPP
#define SAAMGE_B200_PART_HPP

#include "sa_types.hpp"

namespace saamge
{

/// Renumbers parts so every part is connected (amg/src/part.cpp:56-118).
int connectedComponents(std::vector<int> &partitioning, const Table &conn);

/// METIS_PartGraphKway (CONTIG, UFACTOR=30) followed by connectedComponents
/// (amg/src/part.cpp:120-204).  \a weights may be NULL (unit weights).
/// On return *parts holds the actual number of parts.  Caller frees with delete [].
int *part_generate_partitioning(const Table &graph, const int *weights, int *parts);
int *part_generate_partitioning_unweighted(const Table &graph, int *parts);

/// Regular bx x by x bz blocks of a structured grid (deterministic partition
/// independent of METIS; used for fixtures and golden vectors).
int *part_generate_partitioning_blocks(int dim, int nx, int ny, int nz, int bx,
                                       int by, int bz, int *parts);

} // namespace saamge

#endif
