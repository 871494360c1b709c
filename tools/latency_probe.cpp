// Scratch probe (not shipped): per-call latency of gfb_kernel_execute_host for one 47-atom ligand in three 64^3 grids,
// called from C++ (no ctypes overhead).  g++ -O2 -Iinclude tools/latency_probe.cpp -Lopenmmgridforce_b200/lib -lgridforce_b200 -Wl,-rpath,$PWD/openmmgridforce_b200/lib -o /tmp/latency_probe
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "gridforce_b200.h"
#define CHECK(x) do { if ((x) != GFB_OK) { printf("%s: %s\n", #x, gfb_last_error()); return 1; } } while (0)
int main() {
    gfb_device* dev;
    CHECK(gfb_device_open(0, &dev));
    const int counts[3] = {64, 64, 64};
    const double sp[3] = {0.05, 0.05, 0.05}, og[3] = {0, 0, 0};
    std::vector<double> vals(64 * 64 * 64);
    for (size_t i = 0; i < vals.size(); i++) vals[i] = (double) (rand() % 1000) / 100.0;
    gfb_grid* g[3];
    for (int i = 0; i < 3; i++) CHECK(gfb_grid_create(dev, counts, sp, og, vals.data(), vals.size(), GFB_PRECISION_MIXED, GFB_LAYOUT_AUTO, &g[i]));
    const int n = 47;
    std::vector<double> sc(3 * n, 1.0), pos(3 * n), f(3 * n, 0.0);
    for (int i = 0; i < 3 * n; i++) pos[i] = 0.5 + (rand() % 2000) / 1000.0;
    const double k3[3] = {1e4, 1e4, 1e4};
    for (int ng = 3; ng >= 1; ng -= 2) {
        gfb_kernel* k;
        CHECK(gfb_kernel_create(dev, ng, g, n, sc.data(), nullptr, nullptr, k3, &k));
        double e = 0;
        for (int mode = 0; mode < 2; mode++) {
            for (int i = 0; i < 500; i++) CHECK(gfb_kernel_execute_host(k, 1, n, pos.data(), &e, nullptr, f.data(), mode));
            const int reps = 20000;
            auto t0 = std::chrono::steady_clock::now();
            for (int i = 0; i < reps; i++) gfb_kernel_execute_host(k, 1, n, pos.data(), &e, nullptr, f.data(), mode);
            const double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / reps;
            printf("grids=%d mode=%d: %.2f us/call  E=%.6f\n", ng, mode, us, e);
        }
        gfb_kernel_destroy(k);
    }
    return 0;
}
