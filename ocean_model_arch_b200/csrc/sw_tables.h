// sw_tables.h -- columns of the per-row metric tables (no CUDA dependency: also included by the host
// harness that checks the tolerance-mode formulas of sw_fast.cuh on the CPU).
#pragma once

namespace swcu {

// Base table [T_COUNT][h] of doubles, one entry per array row: the nine real(4) metric / Coriolis
// arrays promoted to double, the real(4) sub-expressions of the reference evaluated in real(4) first
// (dx*dy, dy**2, dy/dx ...), and correctly rounded reciprocals (1.0 / value, IEEE division) of the
// divisors the step uses.
enum MetTab : int {
    T_DX, T_DY, T_DXT, T_DYT, T_DXH, T_DYH, T_DXB, T_DYB, T_RLH,
    T_AREA, T_DY2, T_DX2, T_DXB2, T_DYB2, T_RYX, T_RXY, T_RXYB, T_RYXB,
    T_RDXT, T_RDYT, T_RDXH, T_RDYH, T_RDXB, T_RDYB, T_RAREA, T_COUNT
};

}  // namespace swcu
