// spk_comm.h -- the two small collectives of the subtree-partitioned operator evaluation:
// an in-place all-gather of the cut-level entries and a max-all-reduce of the residual pair.
//
// CUDA build: NCCL over NVLink/NVSwitch, one communicator per handle, bootstrapped from a 128-byte ncclUniqueId that
// the host broadcasts (spk_comm_unique_id).  libnccl.so.2 is resolved with dlopen at first use, so single-GPU users
// have no NCCL dependency and a process that already loaded an NCCL (e.g. through torch) shares it.
// Host-emulation test build: the same interface over a POSIX shared-memory segment (ranks = processes or threads).
#pragma once
#include <cstdint>
#include <cstring>
#include <string>

#include "spk_types.h"

#ifndef SPK_HOST_EMU
#include <dlfcn.h>

struct SpkComm {
  int rank = 0, world = 1;
  void *lib = nullptr;
  void *comm = nullptr;
  // minimal NCCL ABI (nccl.h: ncclUniqueId is 128 bytes, ncclFloat64 = 8, ncclMax = 2)
  struct UniqueId { char internal[128]; };
  typedef int (*fn_get_id)(UniqueId *);
  typedef int (*fn_init_rank)(void **, int, UniqueId, int);
  typedef int (*fn_allgather)(const void *, void *, size_t, int, void *, cudaStream_t);
  typedef int (*fn_allreduce)(const void *, void *, size_t, int, int, void *, cudaStream_t);
  typedef int (*fn_destroy)(void *);
  typedef const char *(*fn_errstr)(int);
  fn_get_id get_id = nullptr;
  fn_init_rank init_rank = nullptr;
  fn_allgather allgather = nullptr;
  fn_allreduce allreduce = nullptr;
  fn_destroy destroy_fn = nullptr;
  fn_errstr errstr = nullptr;

  bool load(std::string &err) {
    if (lib) return true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
      lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (lib) break;
    }
    if (!lib) { err = std::string("cannot load NCCL: ") + dlerror(); return false; }
    get_id = (fn_get_id)dlsym(lib, "ncclGetUniqueId");
    init_rank = (fn_init_rank)dlsym(lib, "ncclCommInitRank");
    allgather = (fn_allgather)dlsym(lib, "ncclAllGather");
    allreduce = (fn_allreduce)dlsym(lib, "ncclAllReduce");
    destroy_fn = (fn_destroy)dlsym(lib, "ncclCommDestroy");
    errstr = (fn_errstr)dlsym(lib, "ncclGetErrorString");
    if (!get_id || !init_rank || !allgather || !allreduce || !destroy_fn) { err = "NCCL symbols missing"; return false; }
    return true;
  }
  bool make_id(uint8_t id[128], std::string &err) {
    if (!load(err)) return false;
    UniqueId u;
    const int rc = get_id(&u);
    if (rc != 0) { err = std::string("ncclGetUniqueId: ") + (errstr ? errstr(rc) : "?"); return false; }
    memcpy(id, u.internal, 128);
    return true;
  }
  bool init(int rank_, int world_, const uint8_t id[128], std::string &err) {
    rank = rank_; world = world_;
    if (world <= 1) return true;
    if (!load(err)) return false;
    UniqueId u;
    memcpy(u.internal, id, 128);
    const int rc = init_rank(&comm, world, u, rank);
    if (rc != 0) { err = std::string("ncclCommInitRank: ") + (errstr ? errstr(rc) : "?"); return false; }
    return true;
  }
  // slice g of length count lives at base + g*count on every rank; afterwards every rank holds all slices
  bool allgather_inplace(double *base, long count, spk_stream_t s, std::string &err) {
    if (world <= 1) return true;
    const int rc = allgather(base + (long)rank * count, base, (size_t)count, 8 /*ncclFloat64*/, comm, s);
    if (rc != 0) { err = std::string("ncclAllGather: ") + (errstr ? errstr(rc) : "?"); return false; }
    return true;
  }
  // bytes_per_rank bytes of every rank, rank-major, in place at base (host-side bootstrap data, e.g. IPC handles)
  bool allgather_bytes(void *base, size_t bytes_per_rank, spk_stream_t s, std::string &err) {
    if (world <= 1) return true;
    const int rc = allgather((char *)base + (size_t)rank * bytes_per_rank, base, bytes_per_rank, 1 /*ncclUint8*/, comm, s);
    if (rc != 0) { err = std::string("ncclAllGather: ") + (errstr ? errstr(rc) : "?"); return false; }
    return true;
  }
  bool allreduce_max(double *buf, int n, spk_stream_t s, std::string &err) {
    if (world <= 1) return true;
    // residual maxima are non-negative: the unsigned maximum of their bit patterns is exact and keeps a NaN (spk_max_nn)
    const int rc = allreduce(buf, buf, (size_t)n, 5 /*ncclUint64*/, 2 /*ncclMax*/, comm, s);
    if (rc != 0) { err = std::string("ncclAllReduce: ") + (errstr ? errstr(rc) : "?"); return false; }
    return true;
  }
  bool allreduce_sum(double *buf, int n, spk_stream_t s, std::string &err) {
    if (world <= 1) return true;
    const int rc = allreduce(buf, buf, (size_t)n, 8, 0 /*ncclSum*/, comm, s);
    if (rc != 0) { err = std::string("ncclAllReduce: ") + (errstr ? errstr(rc) : "?"); return false; }
    return true;
  }
  void destroy() {
    if (comm && destroy_fn) destroy_fn(comm);
    comm = nullptr;
  }
};

#else  // ---------------------------------------------------------------- host emulation: shared-memory segment
#include <fcntl.h>
#include <sys/mman.h>
#include <unistd.h>

#include <cstdio>

struct SpkComm {
  int rank = 0, world = 1;
  struct Seg {
    long arrived;
    long generation;
    double data[1 << 17];
  };
  Seg *seg = nullptr;
  char name[64] = {0};

  static void id_to_name(const uint8_t id[128], char *out) {
    static const char *hex = "0123456789abcdef";
    strcpy(out, "/spk_emu_");
    size_t o = strlen(out);
    for (int k = 0; k < 12; ++k) { out[o++] = hex[id[k] >> 4]; out[o++] = hex[id[k] & 15]; }
    out[o] = 0;
  }
  bool make_id(uint8_t id[128], std::string &err) {
    FILE *f = fopen("/dev/urandom", "rb");
    if (!f || fread(id, 1, 128, f) != 128) { err = "cannot read /dev/urandom"; if (f) fclose(f); return false; }
    fclose(f);
    return true;
  }
  bool init(int rank_, int world_, const uint8_t id[128], std::string &err) {
    rank = rank_; world = world_;
    if (world <= 1) return true;
    id_to_name(id, name);
    int fd = -1;
    for (int tries = 0; tries < 20000 && fd < 0; ++tries) {
      fd = shm_open(name, O_RDWR | O_CREAT, 0600);
      if (fd < 0) usleep(1000);
    }
    if (fd < 0) { err = "shm_open failed"; return false; }
    if (ftruncate(fd, sizeof(Seg)) != 0) { err = "ftruncate failed"; close(fd); return false; }
    seg = (Seg *)mmap(nullptr, sizeof(Seg), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (seg == MAP_FAILED) { seg = nullptr; err = "mmap failed"; return false; }
    barrier();
    return true;
  }
  void barrier() {
    const long gen = __atomic_load_n(&seg->generation, __ATOMIC_ACQUIRE);
    if (__atomic_add_fetch(&seg->arrived, 1, __ATOMIC_ACQ_REL) == world) {
      __atomic_store_n(&seg->arrived, 0, __ATOMIC_RELEASE);
      __atomic_add_fetch(&seg->generation, 1, __ATOMIC_ACQ_REL);
    } else {
      while (__atomic_load_n(&seg->generation, __ATOMIC_ACQUIRE) == gen) usleep(50);
    }
  }
  bool allgather_inplace(double *base, long count, spk_stream_t, std::string &err) {
    if (world <= 1) return true;
    if (count * world > (long)(sizeof(seg->data) / sizeof(double))) { err = "emulated all-gather too large"; return false; }
    memcpy(seg->data + (long)rank * count, base + (long)rank * count, sizeof(double) * count);
    barrier();
    memcpy(base, seg->data, sizeof(double) * count * world);
    barrier();
    return true;
  }
  bool allreduce_max(double *buf, int n, spk_stream_t, std::string &) {
    if (world <= 1) return true;
    memcpy(seg->data + (long)rank * n, buf, sizeof(double) * n);
    barrier();
    for (int k = 0; k < n; ++k) {
      double mx = seg->data[k];
      for (int g = 1; g < world; ++g) mx = spk_max_nn(mx, seg->data[(long)g * n + k]);
      buf[k] = mx;
    }
    barrier();
    return true;
  }
  bool allreduce_sum(double *buf, int n, spk_stream_t, std::string &) {  // fixed rank order: deterministic
    if (world <= 1) return true;
    memcpy(seg->data + (long)rank * n, buf, sizeof(double) * n);
    barrier();
    for (int k = 0; k < n; ++k) {
      double sm = 0.0;
      for (int g = 0; g < world; ++g) sm += seg->data[(long)g * n + k];
      buf[k] = sm;
    }
    barrier();
    return true;
  }
  void destroy() {
    if (seg) {
      barrier();
      munmap(seg, sizeof(Seg));
      if (rank == 0) shm_unlink(name);
      seg = nullptr;
    }
  }
};
#endif
