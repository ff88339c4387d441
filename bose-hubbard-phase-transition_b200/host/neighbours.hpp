// neighbours.hpp -- drop-in for the reference's class Neighbours (include/neighbours.hpp:12-63).
// Same constructor and methods; the lists come from the C ABI (bh_neighbours_chain / bh_neighbours_rect).
// Differences, both deliberate: square_neighbours / cube_neighbours fill the member list
// (the reference fills a shadowing local and returns m empty lists), and rect_neighbours is new (any
// lx x ly x lz box, needed for the 4x3 lattice of BASELINE.json config 4).
#pragma once

#include <vector>

class Neighbours {
public:
    Neighbours(int m);
    ~Neighbours();
    void chain_neighbours(bool closed = true);
    void square_neighbours(bool closed = true);  // throws std::invalid_argument unless m is a perfect square
    void cube_neighbours(bool closed = true);    // throws std::invalid_argument unless m is a perfect cube
    void rect_neighbours(int lx, int ly, int lz = 1, bool closed = true);  // requires lx*ly*lz == m
    std::vector<std::vector<int>> getNeighbours() const;

private:
    void fill_box(int lx, int ly, int lz, bool closed);
    int m;
    std::vector<std::vector<int>> neighbours;
};
