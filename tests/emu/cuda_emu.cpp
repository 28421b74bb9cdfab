/*
 * cuda_emu.cpp -- fiber scheduler behind cuda_emu.h (TEST INFRASTRUCTURE).
 * See cuda_emu.h for what this is and is not.
 */
#include "cuda_emu.h"

#include <sys/mman.h>

namespace emu {

thread_local Fiber *cur = nullptr;
thread_local dim3 g_blockDim, g_gridDim;
thread_local std::function<void()> g_entry;

static const size_t kStack = 512 * 1024;

void yield()
{
  Fiber *f = cur;
  swapcontext(&f->ctx, &f->blk->sched);
}

void block_barrier()
{
  Block *b = cur->blk;
  unsigned my = b->gen;
  if(++b->arrived >= b->live) { b->arrived = 0; b->gen++; return; }
  while(b->gen == my) yield();
}

void warp_barrier()
{
  Warp &w = cur->blk->warps[cur->warp];
  unsigned my = w.gen;
  if(++w.arrived >= w.live) { w.arrived = 0; w.gen++; return; }
  while(w.gen == my) yield();
}

uint64_t warp_exchange(uint64_t v, int src_lane)
{
  Fiber *f = cur;
  Warp &w = f->blk->warps[f->warp];
  int buf = f->shfl_count++ & 1;
  w.slot[buf][f->lane] = v;
  warp_barrier();
  return w.slot[buf][src_lane & 31];
}

unsigned warp_ballot(int pred)
{
  Fiber *f = cur;
  Warp &w = f->blk->warps[f->warp];
  int buf = f->shfl_count++ & 1;
  w.slot[buf][f->lane] = pred ? 1 : 0;
  warp_barrier();
  unsigned r = 0;
  int base = f->warp * 32;
  int n = (int)f->blk->fibers.size() - base;
  if(n > 32) n = 32;
  for(int i = 0; i < n; i++)
    if(w.slot[buf][i]) r |= 1u << i;
  return r;
}

static void trampoline()
{
  g_entry();
  Fiber *f = cur;
  Block *b = f->blk;
  f->done = true;
  /* an exited thread no longer takes part in barriers */
  Warp &w = b->warps[f->warp];
  w.live--;
  if(w.live > 0 && w.arrived >= w.live) { w.arrived = 0; w.gen++; }
  b->live--;
  if(b->live > 0 && b->arrived >= b->live) { b->arrived = 0; b->gen++; }
  swapcontext(&f->ctx, &b->sched);
}

void run_grid(dim3 grid, dim3 block, size_t smem)
{
  g_blockDim = block;
  g_gridDim = grid;
  int nt = (int)(block.x * block.y * block.z);
  int nw = (nt + 31) / 32;
  std::vector<void *> stacks(nt);
  for(int t = 0; t < nt; t++) {
    stacks[t] = mmap(nullptr, kStack, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_STACK, -1, 0);
    if(stacks[t] == MAP_FAILED) { perror("emu: mmap"); abort(); }
  }
  unsigned char *dyn = (unsigned char *)aligned_alloc(1024, ((smem + 1023) / 1024 + 1) * 1024);

  for(unsigned bz = 0; bz < grid.z; bz++)
    for(unsigned by = 0; by < grid.y; by++)
      for(unsigned bx = 0; bx < grid.x; bx++) {
        Block blk;
        blk.fibers.resize(nt);
        blk.warps.resize(nw);
        blk.live = nt;
        blk.dyn_smem = dyn;
        blk.bid.x = bx; blk.bid.y = by; blk.bid.z = bz;
        for(int t = 0; t < nt; t++) {
          Fiber &f = blk.fibers[t];
          f.blk = &blk;
          f.tid.x = t % block.x;
          f.tid.y = (t / block.x) % block.y;
          f.tid.z = t / (block.x * block.y);
          f.lane = t & 31;
          f.warp = t >> 5;
          blk.warps[f.warp].live++;
          getcontext(&f.ctx);
          f.ctx.uc_stack.ss_sp = stacks[t];
          f.ctx.uc_stack.ss_size = kStack;
          f.ctx.uc_link = &blk.sched;
          makecontext(&f.ctx, trampoline, 0);
        }
        int remaining = nt;
        while(remaining > 0) {
          remaining = 0;
          for(int t = 0; t < nt; t++) {
            Fiber &f = blk.fibers[t];
            if(f.done) continue;
            cur = &f;
            swapcontext(&blk.sched, &f.ctx);
            if(!f.done) remaining++;
          }
        }
        cur = nullptr;
      }
  free(dyn);
  for(int t = 0; t < nt; t++) munmap(stacks[t], kStack);
}

} // namespace emu
