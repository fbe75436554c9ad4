// Schedule of the streamed host-buffer operator apply (l3b_mf_apply with HOST vectors, the call that stands in for
// Tpetra::Operator::apply → MatrixFreeSystem::applyImpl, algsys/MatrixFreeSystem.hpp:34-41, 1019-1140, when the Krylov vectors live in
// host memory as the reference's do).
//
// The plain form of that call is three serial steps — x over PCIe, the apply, y over PCIe — and at the benchmark's size the two copies
// are 93 % of it (21.3 ms against a 1.4 ms apply, profiles/r2_summary.md). PCIe is full duplex and the element loop touches the vectors
// in a narrow moving window when the elements are ordered the way mesh generators and partitioners order them, so the three steps can
// run as three concurrent streams:
//     copy-in stream   x blocks, in the order the element chunks first need them
//     compute stream   chunk k of the elements as soon as its x blocks have landed; Dirichlet rows of the y blocks that are now final
//     copy-out stream  y blocks whose last contributing chunk has finished
// This file is the host-only, device-free part: cut the vectors into blocks of `block_nodes` nodes, the work into items (element chunks
// and, with a halo, one final item for the border elements + the exchange), record for every block the first and the last item that
// touches it, and emit per item the coalesced node ranges to upload before it and to download after it. Nothing is assumed about the
// numbering: an ordering with no locality degenerates to "everything before item 0, everything after the last item", i.e. to the serial
// form; correctness never depends on the locality.
//
// Invariants (tests/test_host_logic.py): every block is uploaded exactly once, not later than the first item touching it; downloaded
// exactly once, not before the last item touching it, and not before its upload (a block nobody touches still carries Dirichlet rows
// y[d] += alpha x[d]).
#ifndef L3B_APPLY_PLAN_HOST_HPP
#define L3B_APPLY_PLAN_HOST_HPP

#include <algorithm>
#include <cstdint>
#include <stdexcept>
#include <vector>

namespace l3b::host
{
struct ApplyPlan
{
    long long                n_nodes = 0, block_nodes = 0, n_blocks = 0;
    int                      n_items = 0;
    bool                     halo_item = false;  // the last item is "border elements + halo exchange" instead of an element chunk
    std::vector< long long > item_elems;         // [n_items][2]: element range of the item
    std::vector< long long > up_ptr, down_ptr;   // [n_items + 1]: offsets (in ranges) into up_ranges / down_ranges
    std::vector< long long > up_ranges, down_ranges; // [n][2]: node ranges [begin, end)
    long long nUp(int k) const { return up_ptr[k + 1] - up_ptr[k]; }
    long long nDown(int k) const { return down_ptr[k + 1] - down_ptr[k]; }
};

// first / last item touching each block, filled by the touch* helpers
struct BlockTouch
{
    long long          block_nodes = 1;
    std::vector< int > first, last; // first: n_items = untouched so far; last: -1
    BlockTouch(long long n_nodes, long long block_nodes_, int n_items) : block_nodes{block_nodes_}
    {
        if (block_nodes < 1)
            throw std::invalid_argument{"block size must be positive"};
        const auto nb = static_cast< size_t >((n_nodes + block_nodes - 1) / block_nodes);
        first.assign(nb, n_items);
        last.assign(nb, -1);
    }
    void touch(long long node, int item)
    {
        const auto b = static_cast< size_t >(node / block_nodes);
        first[b]     = std::min(first[b], item);
        last[b]      = std::max(last[b], item);
    }
    void touchRange(long long n0, long long n1, int item) // nodes [n0, n1)
    {
        if (n1 <= n0)
            return;
        for (long long b = n0 / block_nodes; b <= (n1 - 1) / block_nodes; ++b)
        {
            first[b] = std::min(first[b], item);
            last[b]  = std::max(last[b], item);
        }
    }
};

// n_border: elements [0, n_border) touch ghost nodes (mesh/SplitMesh.hpp) and wait for the Import; interior elements [n_border, n_elems)
// are cut into chunks of chunk_elems. halo_nodes: the owned nodes this rank sends on Import and receives into on Export. With
// n_border > 0, halo nodes, or ghost nodes (n_owned_nodes < n_nodes) the plan ends with the halo item.
inline ApplyPlan makeApplyPlan(long long n_nodes, long long n_owned_nodes, long long n_elems, int nn, const uint32_t* nodes, long long n_border,
                               const int32_t* halo_nodes, long long n_halo_nodes, long long chunk_elems, long long block_nodes)
{
    if (chunk_elems < 1)
        throw std::invalid_argument{"chunk size must be positive"};
    if (n_border < 0 or n_border > n_elems or n_owned_nodes > n_nodes)
        throw std::invalid_argument{"inconsistent element / node counts"};
    ApplyPlan p;
    p.n_nodes     = n_nodes;
    p.block_nodes = block_nodes;
    p.halo_item   = n_border > 0 or n_halo_nodes > 0 or n_owned_nodes < n_nodes;
    const long long n_interior = n_elems - n_border;
    const int       n_chunks   = static_cast< int >((n_interior + chunk_elems - 1) / chunk_elems);
    p.n_items                  = std::max(1, n_chunks + (p.halo_item ? 1 : 0));
    for (int k = 0; k < n_chunks; ++k)
    {
        p.item_elems.push_back(n_border + k * chunk_elems);
        p.item_elems.push_back(std::min(n_elems, n_border + (k + 1) * chunk_elems));
    }
    if (p.halo_item)
    {
        p.item_elems.push_back(0);
        p.item_elems.push_back(n_border);
    }
    else if (n_chunks == 0) // no elements at all (an empty rank, tests/EmptyPartitionTest.cpp): one empty item carries the copies
    {
        p.item_elems.push_back(0);
        p.item_elems.push_back(0);
    }
    BlockTouch t{n_nodes, block_nodes, p.n_items};
    p.n_blocks = static_cast< long long >(t.first.size());
    for (int k = 0; k < p.n_items; ++k)
        for (long long e = p.item_elems[2 * k]; e < p.item_elems[2 * k + 1]; ++e)
            for (int a = 0; a < nn; ++a)
            {
                const long long node = nodes[e * nn + a];
                if (node >= n_nodes)
                    throw std::out_of_range{"element node id outside the local node range"};
                t.touch(node, k);
            }
    if (p.halo_item)
    {
        const int h = p.n_items - 1;
        for (long long i = 0; i < n_halo_nodes; ++i)
        {
            if (halo_nodes[i] < 0 or halo_nodes[i] >= n_nodes)
                throw std::out_of_range{"halo node id outside the local node range"};
            t.touch(halo_nodes[i], h);
        }
        t.touchRange(n_owned_nodes, n_nodes, h); // the ghost block: written by the Import, read by the Export
    }
    // blocks no item touches travel with item 0
    for (size_t b = 0; b < t.first.size(); ++b)
        if (t.last[b] < 0)
            t.first[b] = t.last[b] = 0;
    const auto emit = [&](const std::vector< int >& when, std::vector< long long >& ptr, std::vector< long long >& ranges) {
        // maximal runs of consecutive blocks with the same item, one pass; then grouped by item (ascending block order within an item)
        std::vector< long long > run_b0, run_b1;
        std::vector< int >       run_item;
        for (long long b = 0; b < p.n_blocks;)
        {
            long long e = b + 1;
            while (e < p.n_blocks and when[static_cast< size_t >(e)] == when[static_cast< size_t >(b)])
                ++e;
            run_b0.push_back(b);
            run_b1.push_back(e);
            run_item.push_back(when[static_cast< size_t >(b)]);
            b = e;
        }
        ptr.assign(p.n_items + 1, 0);
        for (int k : run_item)
            ++ptr[k + 1];
        for (int k = 0; k < p.n_items; ++k)
            ptr[k + 1] += ptr[k];
        ranges.assign(2 * run_item.size(), 0);
        std::vector< long long > fill(ptr.begin(), ptr.end() - 1);
        for (size_t r = 0; r < run_item.size(); ++r)
        {
            const long long at = fill[run_item[r]]++;
            ranges[2 * at]     = run_b0[r] * block_nodes;
            ranges[2 * at + 1] = std::min(n_nodes, run_b1[r] * block_nodes);
        }
    };
    emit(t.first, p.up_ptr, p.up_ranges);
    emit(t.last, p.down_ptr, p.down_ranges);
    return p;
}
} // namespace l3b::host
#endif
