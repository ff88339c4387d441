#include "neighbours.hpp"

#include <cmath>
#include <stdexcept>

#include "../../include/bh_b200.h"

Neighbours::Neighbours(int m_) : m(m_), neighbours(m_ > 0 ? m_ : 0) {}
Neighbours::~Neighbours() {}

static void unpack(int m, const std::vector<int>& ptr, const std::vector<int>& idx, std::vector<std::vector<int>>& out)
{
    out.assign(m, {});
    for (int i = 0; i < m; ++i) out[i].assign(idx.begin() + ptr[i], idx.begin() + ptr[i + 1]);
}

void Neighbours::chain_neighbours(bool closed)
{
    std::vector<int> ptr(m + 1, 0);
    bh_neighbours_chain(m, closed ? 1 : 0, ptr.data(), nullptr);
    std::vector<int> idx(ptr[m] > 0 ? ptr[m] : 1);
    bh_neighbours_chain(m, closed ? 1 : 0, ptr.data(), idx.data());
    // the reference appends to the existing lists (src/neighbours.cpp:21-34); a fresh object is the only use
    std::vector<std::vector<int>> add;
    unpack(m, ptr, idx, add);
    for (int i = 0; i < m; ++i) neighbours[i].insert(neighbours[i].end(), add[i].begin(), add[i].end());
}

void Neighbours::fill_box(int lx, int ly, int lz, bool closed)
{
    std::vector<int> ptr(m + 1, 0);
    bh_neighbours_rect(lx, ly, lz, closed ? 1 : 0, ptr.data(), nullptr);
    std::vector<int> idx(ptr[m] > 0 ? ptr[m] : 1);
    bh_neighbours_rect(lx, ly, lz, closed ? 1 : 0, ptr.data(), idx.data());
    unpack(m, ptr, idx, neighbours);
}

void Neighbours::square_neighbours(bool closed)
{
    const int side = static_cast<int>(std::sqrt(m));
    if (side * side != m) throw std::invalid_argument("The number of sites (m) must be a perfect square.");
    fill_box(side, side, 1, closed);
}

void Neighbours::cube_neighbours(bool closed)
{
    const int side = static_cast<int>(std::cbrt(m));
    if (side * side * side != m) throw std::invalid_argument("The number of sites (m) must be a perfect cube.");
    fill_box(side, side, side, closed);
}

void Neighbours::rect_neighbours(int lx, int ly, int lz, bool closed)
{
    if (lx < 1 || ly < 1 || lz < 1 || lx * ly * lz != m)
        throw std::invalid_argument("The lattice extents must multiply to the number of sites (m).");
    fill_box(lx, ly, lz, closed);
}

std::vector<std::vector<int>> Neighbours::getNeighbours() const { return neighbours; }
