// V3 "OMGRID" binary grid files — the reference's on-disk format (openmmapi/src/GridForce.cpp:495-692 load,
// :694-799 save; openmmapi/src/GridData.cpp:181-267 save with trailer). Host-side, header-only.
//
//   0   char[8]  "OMGRID\0\0"          32  f64 dx,dy,dz          88  u32 grid_type (0 none,1 charge,2 ljr,3 lja)
//   8   u32 version = 3                56  u64 data_offset=128   92  u32 flags
//   12  u32 header_size = 128          64  f64 origin[3]         96  f64 inv_power
//   16  i32 nx,ny,nz                                             104 u32 inv_power_mode (0 NONE,1 RUNTIME,2 STORED)
//   28  u32 deriv_count (0 | 27)                                 108 20 zero bytes
// then nx*ny*nz f64 (x-major, z fastest) — or deriv_count*N f64, derivative-major, function values first.
// GridData::saveToFile appends a trailer: i32 nScaling(=0), f64 origin[3] (and optionally "DERIVS" + data).
#ifndef GF_GRIDFILE_H_
#define GF_GRIDFILE_H_

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>

#include "gridforce_b200.h"

namespace gfb {

struct PackedHeader {   // exactly the 128 bytes on disk (little-endian host assumed, as the reference does)
    char magic[8];
    uint32_t version, header_size;
    int32_t nx, ny, nz;
    uint32_t deriv_count;
    double dx, dy, dz;
    uint64_t data_offset;
    double origin[3];
    uint32_t grid_type, flags;
    double inv_power;
    uint32_t inv_power_mode;
    char reserved[20];
} __attribute__((packed));
static_assert(sizeof(PackedHeader) == 128, "V3 header is 128 bytes");

inline std::string gridfile_read_header(FILE* f, gfb_gridfile_header* out) {
    PackedHeader h;
    if (fread(&h, 1, sizeof h, f) != sizeof h) return "file shorter than the 128-byte V3 header";
    if (memcmp(h.magic, "OMGRID\0\0", 8) != 0) return "Invalid file format (bad magic number)";
    if (h.version != 3) return "Only V3 grid files are supported. Found version " + std::to_string(h.version);
    if (h.nx < 1 || h.ny < 1 || h.nz < 1) return "non-positive grid counts in header";
    if (h.inv_power_mode > 2) return "Invalid inv_power_mode value in file: " + std::to_string(h.inv_power_mode);
    if (h.inv_power_mode != 0 && h.inv_power == 0.0) return "File has inv_power_mode enabled but invalid inv_power value";
    out->counts[0] = h.nx;
    out->counts[1] = h.ny;
    out->counts[2] = h.nz;
    out->spacing[0] = h.dx;
    out->spacing[1] = h.dy;
    out->spacing[2] = h.dz;
    for (int k = 0; k < 3; k++) out->origin[k] = h.origin[k];
    out->grid_type = (int) h.grid_type;
    out->inv_power = h.inv_power;
    out->inv_power_mode = (int) h.inv_power_mode;
    out->deriv_count = h.deriv_count;
    out->data_offset = h.data_offset;
    return std::string();
}

inline void gridfile_fill_header(const gfb_gridfile_header& in, PackedHeader& h) {
    memset(&h, 0, sizeof h);
    memcpy(h.magic, "OMGRID\0\0", 8);
    h.version = 3;
    h.header_size = 128;
    h.nx = in.counts[0];
    h.ny = in.counts[1];
    h.nz = in.counts[2];
    h.deriv_count = 0;
    h.dx = in.spacing[0];
    h.dy = in.spacing[1];
    h.dz = in.spacing[2];
    h.data_offset = 128;
    for (int k = 0; k < 3; k++) h.origin[k] = in.origin[k];
    h.grid_type = (uint32_t) in.grid_type;
    h.flags = 0;
    h.inv_power = in.inv_power;
    h.inv_power_mode = (uint32_t) in.inv_power_mode;
}

}  // namespace gfb
#endif
