// wf_rec.cuh -- packet records exchanged between the kernels of the wave-front pipeline
// (wavefront.cu) and read by the persistent kernel when it finishes the tail (transport.cu).
#pragma once
#include "transport_core.cuh"

namespace mcb {

template <bool MULTI>
__device__ __forceinline__ void rec_store(PacketRec *rec, PacketRecX *recx, const Lane &L, unsigned int pos)
{
    PacketRec r;
    r.rx = L.rx; r.ry = L.ry; r.rz = L.rz; r.passProb = L.passProb;
    r.dx = L.dx; r.dy = L.dy; r.dz = L.dz;
    r.rngn = L.rng.n;
    r.absTau = L.absTau;
    r.istepGen = ((unsigned int)L.istep & 0x7ffffu) | ((unsigned int)L.gen << 19);
    r.k = (unsigned int)L.k;
    r.orgC = L.orgC;
    r.nuP = (unsigned short)L.nuP; r.gP = (unsigned short)L.gP;
    r.flagsLast = (unsigned short)((L.chType & 3) | (L.lgStellar ? 4 : 0) | (L.igpp ? 8 : 0) |
                                   (L.vx != L.dx ? 16 : 0) | (L.vy != L.dy ? 32 : 0) | (L.vz != L.dz ? 64 : 0));
    r.xP = (short)L.xP; r.yP = (short)L.yP; r.zP = (short)L.zP;
    r.orgG = (unsigned short)L.orgG; r.pad = 0;
    const uint4 *src = reinterpret_cast<const uint4 *>(&r);
    uint4 *dst = reinterpret_cast<uint4 *>(&rec[pos]);
    dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2]; dst[3] = src[3];
    if (MULTI) {
        PacketRecX x;
        x.mx = (short)L.mx; x.my = (short)L.my; x.mz = (short)L.mz;
        x.sx = (short)L.sx; x.sy = (short)L.sy; x.sz = (short)L.sz; x.pad = 0;
        *reinterpret_cast<uint4 *>(&recx[pos]) = *reinterpret_cast<const uint4 *>(&x);
    }
}

// FLY kernel: the record is updated in place, so the fields a flight cannot change
// (direction at emission, packet index, origin, generation, nu) are taken from the record
// itself instead of being carried in registers through the cell-crossing loop.
template <bool MULTI>
__device__ __forceinline__ void rec_store_fly(PacketRec *rec, PacketRecX *recx, const Lane &L, unsigned int pos,
                                              unsigned int &k)
{
    uint4 *dst = reinterpret_cast<uint4 *>(&rec[pos]);
    uint4 c1 = dst[1], c2 = dst[2], c3 = dst[3];
    k = c2.z;
    uint4 c0 = make_uint4(__float_as_uint(L.rx), __float_as_uint(L.ry), __float_as_uint(L.rz), __float_as_uint(L.passProb));
    c1.w = L.rng.n;
    c2.x = __float_as_uint(L.absTau);
    c2.y = ((unsigned int)L.istep & 0x7ffffu) | (c2.y & ~0x7ffffu);
    // vHat differs from the stored direction only by mirror-reflection sign flips
    unsigned int flips = (((__float_as_uint(L.vx) ^ c1.x) >> 31) << 4) | (((__float_as_uint(L.vy) ^ c1.y) >> 31) << 5) |
                         (((__float_as_uint(L.vz) ^ c1.z) >> 31) << 6);
    unsigned int flags = (unsigned int)(L.chType & 3) | (L.lgStellar ? 4u : 0u) | (L.igpp ? 8u : 0u) | flips;
    c3.x = (c3.x & 0xffffu) | ((unsigned int)(unsigned short)L.gP << 16);
    c3.y = flags | ((unsigned int)(unsigned short)(short)L.xP << 16);
    c3.z = (unsigned int)(unsigned short)(short)L.yP | ((unsigned int)(unsigned short)(short)L.zP << 16);
    dst[0] = c0; dst[1] = c1; dst[2] = c2; dst[3] = c3;
    if (MULTI) {
        PacketRecX x;
        x.mx = (short)L.mx; x.my = (short)L.my; x.mz = (short)L.mz;
        x.sx = (short)L.sx; x.sy = (short)L.sy; x.sz = (short)L.sz; x.pad = 0;
        *reinterpret_cast<uint4 *>(&recx[pos]) = *reinterpret_cast<const uint4 *>(&x);
    }
}

template <bool MULTI>
__device__ __forceinline__ void rec_load(const TransportArgs &t, const PacketRec *rec, const PacketRecX *recx,
                                         Lane &L, unsigned int pos)
{
    PacketRec r;
    const uint4 *src = reinterpret_cast<const uint4 *>(&rec[pos]);
    uint4 *dst = reinterpret_cast<uint4 *>(&r);
    dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2]; dst[3] = src[3];
    L.k = (long long)r.k;
    L.rng.init(t.seed, packet_pid(t, (long long)r.k), t.rngStream, r.rngn);
    L.rx = r.rx; L.ry = r.ry; L.rz = r.rz; L.passProb = r.passProb;
    L.dx = r.dx; L.dy = r.dy; L.dz = r.dz;
    // vHat = direction, except for the signs a mirror reflection flipped (continued flights)
    L.vx = (r.flagsLast & 16) ? -r.dx : r.dx;
    L.vy = (r.flagsLast & 32) ? -r.dy : r.dy;
    L.vz = (r.flagsLast & 64) ? -r.dz : r.dz;
    set_inv(L);
    L.absTau = r.absTau;
    L.segs = 0; L.istep = (int)(r.istepGen & 0x7ffffu); L.gen = (int)(r.istepGen >> 19);
    L.nuP = r.nuP; L.gP = r.gP;
    L.chType = r.flagsLast & 3; L.lgStellar = (r.flagsLast >> 2) & 1; L.igpp = (r.flagsLast >> 3) & 1;
    L.lastNuP = r.nuP;                       // a stored packet's last emission is its current nu
    L.xP = r.xP; L.yP = r.yP; L.zP = r.zP;
    L.orgG = r.orgG; L.orgC = r.orgC;
    L.fate = 0; L.pendFate = FATE_ESCAPED; L.planeG = 0;
    if (MULTI) {
        PacketRecX x;
        *reinterpret_cast<uint4 *>(&x) = *reinterpret_cast<const uint4 *>(&recx[pos]);
        L.mx = x.mx; L.my = x.my; L.mz = x.mz; L.sx = x.sx; L.sy = x.sy; L.sz = x.sz;
    } else {
        L.mx = L.xP; L.my = L.yP; L.mz = L.zP;       // single grid: the mother slot is the cell
        L.sx = L.sy = L.sz = -1;
    }
}


}  // namespace mcb
