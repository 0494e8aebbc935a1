// Host emulation of namespace kb200::tma (included by kornia_b200/csrc/warp_tma.cuh only when KB200_HOST_EMU is
// defined, i.e. never in the product build).  The mbarrier lives in the kernel's own shared-memory word; TMA operations
// go through the scheduler of hostemu.h, which completes them either at once ("eager": exposes loads that overwrite
// data still in use) or only when every thread is blocked ("lazy": exposes reads that do not wait for their load).
struct EmuMap {            // what the emulator keeps inside the 128 opaque bytes of a CUtensorMap
  const float* base;
  uint64_t dims[3];        // elements, innermost first
  uint64_t strides[2];     // bytes
  uint32_t box[3];
};
struct EmuBar {            // the 8 bytes of an mbarrier
  uint32_t phase;          // completed phases
  int16_t pending;         // arrivals still missing in the current phase
  int16_t count;           // arrivals per phase
};
void emu_tma_issue(void* dst, const EmuMap* map, uint64_t* bar, int c0, int c1, int c2);  // hostemu.h
void emu_spin();                                                                            // yield while polling
long long& emu_bar_tx(uint64_t* bar);                                                       // outstanding bytes of a barrier

// 32-bit shared-memory addresses: offsets into the CTA's shared array (emu_set_smem), so that the kernels' modular
// address arithmetic works and every ld/st.shared is bounds-checked
uint32_t emu_smem_addr(const void* p);
float* emu_smem_ptr(uint32_t addr);
inline uint32_t smem_u32(const void* p) { return emu_smem_addr(p); }
inline void mbar_init(uint64_t* bar, uint32_t count) {
  EmuBar b{0u, (int16_t)count, (int16_t)count};
  memcpy(bar, &b, 8);
  emu_bar_tx(bar) = 0;
}
inline void fence_barrier_init() {}
inline void fence_proxy_async() {}
inline void emu_bar_try_complete(uint64_t* bar) {
  EmuBar* b = reinterpret_cast<EmuBar*>(bar);
  if (b->pending == 0 && emu_bar_tx(bar) == 0) {
    ++b->phase;
    b->pending = b->count;
  }
}
inline void mbar_arrive(uint64_t* bar) {
  --reinterpret_cast<EmuBar*>(bar)->pending;
  emu_bar_try_complete(bar);
}
inline void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  emu_bar_tx(bar) += bytes;
  --reinterpret_cast<EmuBar*>(bar)->pending;
  emu_bar_try_complete(bar);
}
inline bool mbar_try_wait(uint64_t* bar, uint32_t parity) { return (reinterpret_cast<EmuBar*>(bar)->phase & 1u) != parity; }
inline void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) emu_spin();
}
inline void load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  emu_tma_issue(dst, reinterpret_cast<const EmuMap*>(map), bar, c0, c1, c2);
}
long long& emu_lds_count();
inline float lds(uint32_t a) {
  ++emu_lds_count();
  return *emu_smem_ptr(a);
}
inline float lds_ro(uint32_t a) { return *emu_smem_ptr(a); }
inline void sts(uint32_t a, float v) { *emu_smem_ptr(a) = v; }
inline bool elect_one() {  // elect.sync: a warp collective that picks one lane
  emu_warp_exchange(0ull, emu_lane());
  return emu_lane() == 0;
}
void emu_reduce_add_issue(const EmuMap* map, uint32_t src_smem, int c0, int c1, int c2);
void emu_bulk_wait(bool read_only);
inline void reduce_add_3d(const CUtensorMap* map, uint32_t src_smem, int c0, int c1, int c2) {
  emu_reduce_add_issue(reinterpret_cast<const EmuMap*>(map), src_smem, c0, c1, c2);
}
inline void bulk_commit() {}
inline void bulk_wait_read0() { emu_bulk_wait(true); }
inline void bulk_wait0() { emu_bulk_wait(false); }
inline float rcp_approx(float den) { return 1.0f / den; }  // any approximation within the refinement's basin gives the same quotients
inline void prefetch_3d(const CUtensorMap*, int, int, int) {}
inline void prefetch_map(const CUtensorMap*) {}
inline void prefetch_l1(size_t) {}
