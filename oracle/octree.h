// ORACLE -- test infrastructure, not product code.
//
// CPU restatement of the SAMPLER half of fidget-mesh's Manifold Dual
// Contouring octree build:
//   OctreeBuilder::recurse   <-> fidget-mesh/src/octree.rs:521-583
//   OctreeBuilder::leaf      <-> fidget-mesh/src/octree.rs:590-808 (corner
//                                samples, 4 rounds of 16-ary edge search,
//                                gradient at each intersection)
//   CellBounds::{child,corner,pos}, DirectedEdge::to_undirected
//                            <-> fidget-mesh/src/cell.rs:155-196, types.rs:208-219
// Out of scope: QEF solve, cell collapse, walk_dual, STL.
// The set of active edges of a leaf is determined by its corner mask (an
// edge is active iff its two corners differ in sign), so the per-vertex
// grouping table CELL_TO_VERT_TO_EDGES (build.rs) is not needed to reproduce
// LeafHermiteData.intersections, which is indexed by undirected edge.
#pragma once
#include <vector>

#include "vm.h"

namespace oracle {

struct OctreeLeaf {
    uint16_t ix, iy, iz;   // cell coordinates at the maximum depth
    uint8_t mask;          // bit i set = corner i inside (bit 0 = +X, 1 = +Y, 2 = +Z)
    uint8_t n_edges;
    uint16_t present;      // bit e set = undirected edge e carries an intersection
    uint16_t pad;
    float pos[12][3];      // LeafIntersection::pos.xyz, indexed by undirected edge
    float grad[12][4];     // LeafIntersection::grad = (dx, dy, dz, v)
};

struct OctreeStats {
    uint64_t evaluated[16] = {0}, full[16] = {0}, empty[16] = {0}, ambiguous[16] = {0};
    uint64_t leaf_empty = 0, leaf_full = 0, leaf_surface = 0;
    uint64_t float_points = 0, grad_points = 0;
};

struct OctreeConfig {
    uint32_t depth = 0;
    bool has_transform = false;   // Settings::world_to_model != identity (octree.rs:493-498)
    Mat4 world_to_model;
    int threads = 1;
};

void octree_sample(const TapeP& tape, const OctreeConfig& cfg, std::vector<OctreeLeaf>& leaves, OctreeStats* stats);

}  // namespace oracle
