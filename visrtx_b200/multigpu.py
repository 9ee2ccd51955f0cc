"""Multi-GPU drivers for the DVR path on one NVSwitch box: one process per GPU (torch.distributed is
plumbing only — rendezvous, handle exchange, stream-ordered barriers).

* sort-first  (volume fits one GPU; BASELINE C2/C3): every rank holds the whole field and renders the
  tile rows ``row % world == rank``.  The render kernel's colour stores go straight into the display
  rank's frame through a CUDA-IPC mapped pointer (peer stores over NVLink), so the "gather" is fused
  into the render launch; accumulation/depth stay sharded with their owner.
* sort-last   (C4): every rank holds a z-slab (+1 ghost slice each side) and marches its samples of
  the GLOBAL lattice into a premultiplied partial image; then ONE fused kernel per rank composites its
  pixel strip from all ranks' partial images (peer loads over NVLink, per-pixel view order) and
  resolves it into the display rank's frame.  Partial images are double-buffered so one barrier per
  frame separates "all partials written" from "composite".

The partition helpers are pure Python and tested with gloo on CPU (tests/test_multigpu_host.py).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple


def slab_ranges(nz: int, world: int) -> List[Tuple[int, int]]:
    """Balanced, contiguous z-slice ownership [z0, z1) per rank; every slice owned exactly once."""
    if world < 1 or nz < world:
        raise ValueError(f"cannot split {nz} slices over {world} ranks")
    base, rem = divmod(nz, world)
    out, z = [], 0
    for r in range(world):
        n = base + (1 if r < rem else 0)
        out.append((z, z + n))
        z += n
    return out


def view_balanced_slab_ranges(nz: int, world: int, bounds_lo, bounds_hi, eye=None, min_slices: int = 2
                              ) -> List[Tuple[int, int]]:
    """z-slice ownership balanced by WORK for a perspective camera at `eye` (object space).

    Lattice samples are uniform along each ray, and rays are uniform per solid angle, so the sample density in
    space falls off as 1/r^2 from the eye: the slab nearest the camera holds up to (r_far/r_near)^2 times the samples
    of the farthest one (1.76x at the 2|diag| benchmark orbit).  Slices are weighted by the mean 1/r^2 over their
    area and the boundaries are put at equal cumulative weight.  eye=None (orthographic) gives `slab_ranges`.
    The partition is static: it suits a camera that stays near the pose it was computed for."""
    if eye is None:
        return slab_ranges(nz, world)
    if world < 1 or nz < world * min_slices:
        raise ValueError(f"cannot split {nz} slices over {world} ranks with >= {min_slices} slices each")
    lo = [float(v) for v in bounds_lo]
    hi = [float(v) for v in bounds_hi]
    ex, ey, ez = (float(v) for v in eye)
    g = 12
    xs = [lo[0] + (hi[0] - lo[0]) * (i + 0.5) / g for i in range(g)]
    ys = [lo[1] + (hi[1] - lo[1]) * (j + 0.5) / g for j in range(g)]
    w = []
    for k in range(nz):
        z = lo[2] + (hi[2] - lo[2]) * (k + 0.5) / nz
        acc = 0.0
        for x in xs:
            for y in ys:
                r2 = (x - ex) ** 2 + (y - ey) ** 2 + (z - ez) ** 2
                acc += 1.0 / max(r2, 1e-12)
        w.append(acc)
    total = sum(w)
    cuts, run, target = [0], 0.0, 1
    for k in range(nz):
        run += w[k]
        while target < world and run >= total * target / world:
            cuts.append(k + 1)
            target += 1
    cuts = cuts[:world] + [nz]
    # enforce the minimum thickness front to back, then back to front
    for i in range(1, world):
        cuts[i] = max(cuts[i], cuts[i - 1] + min_slices)
    for i in range(world - 1, 0, -1):
        cuts[i] = min(cuts[i], cuts[i + 1] - min_slices)
    return [(cuts[i], cuts[i + 1]) for i in range(world)]


def slab_margin(nz: int, world: int, cap: int = 64) -> int:
    """Extra slices a slab keeps resident on each side of its initial range so that ownership can move later
    (dvr_field_set_owned_slices): half of the mean slab thickness, between 4 and `cap` slices.  (A quarter was not
    enough at N = 8: the cuts of the time-balanced partition sit up to 30 slices from the 1/r^2 model's, and the rounds
    that balance the fused frame's march phases need room beyond that.)"""
    return max(4, min(cap, nz // (2 * max(world, 1))))


def creation_ranges(ranges: Sequence[Tuple[int, int]], nz: int, margin: int) -> List[Tuple[int, int]]:
    """The ownership LIMITS every slab is created with: its initial range widened by `margin` slices on each side."""
    return [(max(z0 - margin, 0), min(z1 + margin, nz)) for z0, z1 in ranges]


def rebalance_slab_ranges(ranges: Sequence[Tuple[int, int]], times: Sequence[float],
                          limits: Sequence[Tuple[int, int]], min_slices: int = 2,
                          damping: float = 1.0) -> List[Tuple[int, int]]:
    """New z-slice ownership from MEASURED per-GPU march times (any unit) of the current partition.

    The cost of a slice is taken as uniform inside a slab (time / thickness); the cuts go where the cumulative cost
    reaches k/N of the total, clamped to what both neighbours hold resident (`limits` = the ranges the slabs were
    created with, see creation_ranges).  No voxel moves: every GPU applies its new range with
    dvr_field_set_owned_slices before the next frame.  Deterministic, so ranks that all-gather the same times compute
    the same cuts.  Apply repeatedly (2-3 rounds) — the cost density inside a slab is not really uniform.
    `damping` < 1 moves every cut only that fraction of the way (for times that contain a part which does not scale
    with the slab's thickness, e.g. the march phase measured inside the fused frame)."""
    world = len(ranges)
    if world != len(times) or world != len(limits):
        raise ValueError("ranges, times and limits must have one entry per rank")
    nz = ranges[-1][1]
    if world == 1:
        return [tuple(ranges[0])]
    total = float(sum(times))
    if not total > 0.0:
        return [tuple(r) for r in ranges]
    cuts = [ranges[0][0]]
    acc, r = 0.0, 0
    for k in range(1, world):
        target = total * k / world
        while r < world - 1 and acc + times[r] < target:
            acc += times[r]
            r += 1
        z0, z1 = ranges[r]
        frac = (target - acc) / times[r] if times[r] > 0 else 0.0
        cuts.append(int(round(z0 + frac * (z1 - z0))))
    cuts.append(nz)
    if damping < 1.0:
        for i in range(1, world):
            old = ranges[i][0]
            cuts[i] = int(round(old + damping * (cuts[i] - old)))
    for i in range(1, world):  # what the two neighbours of cut i can own, and a minimum thickness front to back
        lo = max(limits[i][0], cuts[i - 1] + min_slices)
        hi = limits[i - 1][1]
        cuts[i] = min(max(cuts[i], lo), hi)
    for i in range(world - 1, 0, -1):
        cuts[i] = min(cuts[i], cuts[i + 1] - min_slices)
    out = [(cuts[i], cuts[i + 1]) for i in range(world)]
    for (z0, z1), (l0, l1) in zip(out, limits):
        if z0 < l0 or z1 > l1 or z1 - z0 < 1:
            return [tuple(r) for r in ranges]  # the limits leave no valid partition: keep the current one
    return out


def resident_range(z0: int, z1: int, nz: int) -> Tuple[int, int]:
    """Slices a slab must hold: its own plus one ghost slice on each side, clamped to the volume."""
    return max(z0 - 1, 0), min(z1 + 1, nz)


def pixel_strips(n_pixels: int, world: int, align: int = 256) -> List[Tuple[int, int]]:
    """Contiguous pixel ranges for the compositing stage, aligned so that stores stay coalesced."""
    per = -(-n_pixels // world)
    per = -(-per // align) * align
    out = []
    for r in range(world):
        b = min(r * per, n_pixels)
        out.append((b, min(b + per, n_pixels)))
    return out


def tile_rows_of(rank: int, world: int, height: int, tile_h: int = 4) -> List[int]:
    """Tile rows (of tile_h pixel rows) a rank renders in sort-first mode."""
    n_rows = -(-height // tile_h)
    return [r for r in range(n_rows) if r % world == rank]


def exchange_bytes(dist, payload: bytes, world: int, group=None) -> List[bytes]:
    """all-gather of one small bytes object per rank (CUDA-IPC handles); works on gloo and nccl."""
    out = [None] * world
    dist.all_gather_object(out, payload, group=group)
    return out


class SharedHostFrame:
    """One host copy of the display frame that EVERY rank's GPU can store into: a POSIX shared-memory segment
    created by rank 0, mapped by all ranks and pinned + device-mapped in each process (cudaHostRegister).  Its
    device address goes into DvrFrameBuffers::outColorMirror, so each rank's resolve / render kernel streams its
    pixels of the final image to the host during the launch — the multi-process version of the single-GPU host
    streaming; no device->host copy after the frame."""

    def __init__(self, dist, rank: int, world: int, nbytes: int, group=None):
        import ctypes
        from multiprocessing import shared_memory
        self.rank, self.nbytes, self.group = rank, nbytes, group
        name = ""
        if rank == 0:
            self.shm = shared_memory.SharedMemory(create=True, size=nbytes)
            name = self.shm.name
        names = exchange_bytes(dist, name.encode(), world, group) if world > 1 else [name.encode()]
        if rank != 0:
            self.shm = shared_memory.SharedMemory(name=names[0].decode())
            try:  # the creator unlinks the segment; attachments must not be tracked (Python < 3.13 has no track=False)
                from multiprocessing import resource_tracker
                resource_tracker.unregister(self.shm._name, "shared_memory")
            except Exception:
                pass
        self._cudart = ctypes.CDLL("libcudart.so")
        self.host_ptr = ctypes.addressof(ctypes.c_char.from_buffer(self.shm.buf))
        # cudaHostRegisterPortable | cudaHostRegisterMapped
        rc = self._cudart.cudaHostRegister(ctypes.c_void_p(self.host_ptr), ctypes.c_size_t(nbytes), ctypes.c_uint(1 | 2))
        if rc != 0:
            raise RuntimeError(f"cudaHostRegister failed ({rc})")
        dp = ctypes.c_void_p()
        rc = self._cudart.cudaHostGetDevicePointer(ctypes.byref(dp), ctypes.c_void_p(self.host_ptr), ctypes.c_uint(0))
        if rc != 0:
            raise RuntimeError(f"cudaHostGetDevicePointer failed ({rc})")
        self.dev_ptr = dp.value

    def numpy(self, dtype="uint32"):
        import numpy as np
        return np.frombuffer(self.shm.buf, dtype=dtype, count=self.nbytes // np.dtype(dtype).itemsize)

    def close(self, dist=None):
        import ctypes
        self._cudart.cudaHostUnregister(ctypes.c_void_p(self.host_ptr))
        if dist is not None:
            dist.barrier(group=self.group)
        try:
            self.shm.close()
        except BufferError:  # a numpy view is still alive; the segment goes away with the process
            pass
        if self.rank == 0:
            try:
                self.shm.unlink()
            except FileNotFoundError:
                pass


class _Barrier:
    """Stream-ordered barrier: a 1-element all-reduce enqueued on the current stream (no host sync)."""

    def __init__(self, dist, torch, device):
        self.dist = dist
        self.flag = torch.zeros(1, dtype=torch.int32, device=device)

    def __call__(self):
        self.dist.all_reduce(self.flag)


class SortFirst:
    """Sort-first driver.  `render(frame_id, camera)` enqueues one frame on the current stream."""

    def __init__(self, capi, torch, dist, rank: int, world: int, device, width: int, height: int, instances,
                 n_instances: int, fmt: int, integrator: int, rate: float, background, skip: bool = False,
                 tile_band: int = 1, host_mirror: bool = False):
        self.capi, self.torch, self.dist = capi, torch, dist
        self.rank, self.world, self.device = rank, world, device
        self.W, self.H, self.fmt = width, height, fmt
        self.integrator, self.rate, self.background, self.skip = integrator, rate, background, skip
        self.instances, self.n_instances = instances, n_instances
        self.tile_band = tile_band
        npx = width * height
        px_bytes = 16 if fmt == capi.DVR_FORMAT_FLOAT32_VEC4 else 4
        self.accum = torch.zeros((npx, 4), dtype=torch.float32, device=device)
        self.depth = torch.zeros(npx, dtype=torch.float32, device=device)
        self._owned = None
        if world == 1:
            self.color_local = torch.zeros(npx * px_bytes // 4, dtype=torch.int32, device=device)
            self.color_ptr = self.color_local.data_ptr()
        else:
            if rank == 0:
                self._owned, handle = capi.ipc_alloc(npx * px_bytes)
                self.color_ptr = self._owned
            else:
                handle = b""
            handles = exchange_bytes(dist, handle, world)
            if rank != 0:
                self.color_ptr = capi.ipc_open(handles[0])
        self.host_frame = SharedHostFrame(dist, rank, world, npx * px_bytes) if host_mirror else None
        self._fb_plain = capi.frame_buffers(self.accum.data_ptr(), self.color_ptr, self.depth.data_ptr())
        self._fb_mirror = capi.frame_buffers(self.accum.data_ptr(), self.color_ptr, self.depth.data_ptr(),
                                             color_mirror=self.host_frame.dev_ptr) if self.host_frame else None
        self.fb = self._fb_mirror or self._fb_plain
        self.barrier = _Barrier(dist, torch, device) if world > 1 else (lambda: None)

    def params(self, frame_id: int):
        return self.capi.frame_params(self.W, self.H, self.fmt, self.integrator, frame_id, -1, 1, self.rate,
                                      self.background, tile_rank=self.rank, tile_ranks=self.world, skip=self.skip,
                                      tile_band=self.tile_band)

    def stream_to_host(self, on: bool):
        """Route the final colour also into the shared host frame (needs host_mirror=True at construction)."""
        self.fb = self._fb_mirror if (on and self._fb_mirror is not None) else self._fb_plain

    def render(self, frame_id: int, camera, stream: int):
        self.capi.render(self.params(frame_id), camera, self.instances, self.n_instances, self.fb, stream)
        self.barrier()  # all ranks' tile rows have landed in the display rank's frame

    def color_tensor(self):
        """Display rank only: the assembled frame as a torch tensor view (copy)."""
        import ctypes
        npx = self.W * self.H
        n32 = npx * (4 if self.fmt == self.capi.DVR_FORMAT_FLOAT32_VEC4 else 1)
        out = self.torch.empty(n32, dtype=self.torch.int32, device=self.device)
        ctypes.CDLL("libcudart.so").cudaMemcpy(ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(self.color_ptr),
                                               ctypes.c_size_t(n32 * 4), ctypes.c_int(3))
        return out

    def close(self):
        self.torch.cuda.synchronize()
        if self.host_frame:
            self.host_frame.close(self.dist if self.world > 1 else None)
        if self.world > 1:
            self.dist.barrier()
            if self.rank != 0:
                self.capi.ipc_close(self.color_ptr)
            self.dist.barrier()
            if self._owned:
                self.capi.ipc_free(self._owned)


class SortLast:
    """Sort-last driver over z-slabs: ONE fused launch per GPU and frame (dvr_render_slab_frame) — march of the own
    slab, per-region completion flags to the regions' owners, compositing + resolve of the owned regions over peer
    memory as soon as their inputs are complete, background strips.

    Cross-GPU ordering is done ON THE DEVICE with flags in CUDA-IPC shared memory: region flags (rank q finished
    region r of frame seq) and resolved flags (rank q finished compositing frame seq: partial buffers alternate, and
    the display rank's launch completes only when every rank's pixels have landed).  No host synchronisation, no
    NCCL call and no second launch on the per-frame path.  `fused=False` keeps the round-1 sequence (partial march,
    wait, peer composite, signal: three launches after the march) for A/B measurements.
    """

    MAX_REGIONS = 1024
    # per-rank flag table (uint32): [MAX_REGIONS][16] region flags, [16] resolved flags, [16] error word + padding,
    # then the round-1 layout ([0:16] partial-complete, [16:32] strip-resolved, [32] error) for fused=False
    FLAG_WORDS = MAX_REGIONS * 16 + 32 + 64

    def __init__(self, capi, torch, dist, rank: int, world: int, device, width: int, height: int, instance,
                 obj_id: int, inst_id: int, fmt: int, integrator: int, rate: float, background, skip: bool = False,
                 host_mirror: bool = False, fused: bool = True, group=None):
        self.capi, self.torch, self.dist = capi, torch, dist
        self.fused, self.group = fused, group
        self.rank, self.world, self.device = rank, world, device
        self.W, self.H, self.fmt = width, height, fmt
        self.integrator, self.rate, self.background, self.skip = integrator, rate, background, skip
        self.instance, self.obj_id, self.inst_id = instance, obj_id, inst_id
        npx = width * height
        self.npx = npx
        px_bytes = 16 if fmt == capi.DVR_FORMAT_FLOAT32_VEC4 else 4
        self.strips = pixel_strips(npx, world)
        self.accum = torch.zeros((npx, 4), dtype=torch.float32, device=device)
        self.depth = torch.zeros(npx, dtype=torch.float32, device=device)
        self._mine, self._peer_open = [], []
        self._color_owned = None
        self.frames = 0
        if world == 1:
            self.buf = [torch.zeros((npx, 5), dtype=torch.float32, device=device) for _ in range(2)]
            self.rgba_ptrs = [[b.data_ptr()] for b in self.buf]
            self.depth_ptrs = [[b.data_ptr() + npx * 16] for b in self.buf]
            self.color_local = torch.zeros(npx * px_bytes // 4, dtype=torch.int32, device=device)
            self.color_ptr = self.color_local.data_ptr()
            self.flag_ptrs = None
        else:
            payload = b""
            for _ in range(2):
                p, h = capi.ipc_alloc(npx * 20)
                self._mine.append(p)
                payload += h
            fp, fh = capi.ipc_alloc(self.FLAG_WORDS * 4)
            self._mine.append(fp)
            import ctypes
            ctypes.CDLL("libcudart.so").cudaMemset(ctypes.c_void_p(fp), 0, ctypes.c_size_t(self.FLAG_WORDS * 4))
            torch.cuda.synchronize()
            payload += fh
            if rank == 0:  # the display frame: colour, then depth (assembled on the display GPU like the colour)
                self._color_owned, hc = capi.ipc_alloc(npx * px_bytes + npx * 4)
                payload += hc
            all_h = exchange_bytes(dist, payload, world, group)
            self.rgba_ptrs, self.depth_ptrs, self.flag_ptrs = [[], []], [[], []], []
            for r in range(world):
                for b in range(2):
                    if r == rank:
                        p = self._mine[b]
                    else:
                        p = capi.ipc_open(all_h[r][64 * b:64 * (b + 1)])
                        self._peer_open.append(p)
                    self.rgba_ptrs[b].append(p)
                    self.depth_ptrs[b].append(p + npx * 16)
                if r == rank:
                    self.flag_ptrs.append(fp)
                else:
                    q = capi.ipc_open(all_h[r][128:192])
                    self._peer_open.append(q)
                    self.flag_ptrs.append(q)
            if rank == 0:
                self.color_ptr = self._color_owned
            else:
                self.color_ptr = capi.ipc_open(all_h[0][192:256])
                self._peer_open.append(self.color_ptr)
            dist.barrier(group=group)  # every table is zeroed and mapped before the first signal can arrive
            self.region_done = torch.zeros(self.MAX_REGIONS, dtype=torch.int32, device=device)
        self.host_frame = SharedHostFrame(dist, rank, world, npx * px_bytes, group) if host_mirror else None
        # world > 1: depth is assembled in the display rank's frame next to the colour.  The display rank resolves
        # straight into it; the others keep the depth of their own pixels locally (the accumulate step reads it back)
        # and mirror every store into the display frame through the peer pointer — write-only traffic over NVLink
        self.display_depth_ptr = self.depth.data_ptr() if world == 1 else self.color_ptr + npx * px_bytes
        depth_ptr = self.display_depth_ptr if rank == 0 else self.depth.data_ptr()
        depth_mirror = 0 if rank == 0 else self.display_depth_ptr
        self._fb_plain = capi.frame_buffers(self.accum.data_ptr(), self.color_ptr, depth_ptr, depth_mirror=depth_mirror)
        self._fb_mirror = capi.frame_buffers(self.accum.data_ptr(), self.color_ptr, depth_ptr, depth_mirror=depth_mirror,
                                             color_mirror=self.host_frame.dev_ptr) if self.host_frame else None
        self.fb = self._fb_mirror or self._fb_plain
        self.frame_parity = 0
        self._hot = None
        self.timing_ptr = 0  # fused launches stamp %globaltimer phases into this uint64[8] when set (bench bookkeeping)

    def params(self, frame_id: int):
        # the synchronised fused composite (world > 1) regenerates every primary ray and never reads a pixel whose ray
        # misses the bounds, so the partial march may confine itself to the screen rectangle of the volume
        return self.capi.frame_params(self.W, self.H, self.fmt, self.integrator, frame_id, -1, 1, self.rate,
                                      self.background, skip=self.skip, partial_cull_to_bounds=self.world > 1)

    def stream_to_host(self, on: bool):
        """Route the final colour also into the shared host frame (needs host_mirror=True at construction)."""
        self.fb = self._fb_mirror if (on and self._fb_mirror is not None) else self._fb_plain

    def render(self, frame_id: int, camera, stream: int, wait_display: bool = True):
        capi = self.capi
        b = self.frame_parity
        self.frame_parity ^= 1
        self.frames += 1
        seq = self.frames  # monotonically increasing frame number carried by the flags
        p = self.params(frame_id)
        lo, hi = self.strips[self.rank]
        if self.world == 1:
            capi.render_partial(p, camera, self.instance, self.rgba_ptrs[b][0], self.depth_ptrs[b][0], stream)
            capi.composite_resolve_peers(p, camera, self.rgba_ptrs[b], self.depth_ptrs[b], self.obj_id, self.inst_id,
                                         self.fb, lo, hi, stream)
            return
        me, W = self.rank, self.world
        if self.fused:
            return self._render_fused(frame_id, camera, stream, b, seq, wait_display)
        leg = (self.MAX_REGIONS * 16 + 32) * 4  # the round-1 flag layout lives behind the fused tables
        my_flags = self.flag_ptrs[me] + leg
        err = self.flag_ptrs[me] + (self.MAX_REGIONS * 16 + 16) * 4
        if self._hot is None:  # per-frame host work is pointer/struct reuse only: everything below is built once
            import ctypes as C
            L = capi.lib
            self._hot = {
                "p": self.params(0),
                "sync_p": capi.peer_sync(signal_ptrs=[self.flag_ptrs[r] + leg + me * 4 for r in range(W)], signal_value=0),
                "sync_c": capi.peer_sync(signal_ptrs=[self.flag_ptrs[r] + leg + (16 + me) * 4 for r in range(W)],
                                         signal_value=0, wait_ptr=my_flags, n_wait=W, wait_value=0, error_flag=err),
                "rg": [(C.c_void_p * W)(*self.rgba_ptrs[k]) for k in range(2)],
                "dp": [(C.c_void_p * W)(*self.depth_ptrs[k]) for k in range(2)],
                "mine_rg": [C.c_void_p(self.rgba_ptrs[k][me]) for k in range(2)],
                "mine_dp": [C.c_void_p(self.depth_ptrs[k][me]) for k in range(2)],
                "resolved": C.c_void_p(my_flags + 16 * 4), "err": C.c_void_p(err), "W": C.c_uint32(W),
                "obj": C.c_uint32(self.obj_id), "inst": C.c_uint32(self.inst_id),
                "lo": C.c_size_t(lo), "hi": C.c_size_t(hi), "C": C, "L": L,
            }
        h = self._hot
        C, L = h["C"], h["L"]
        p, sync_p, sync_c = h["p"], h["sync_p"], h["sync_c"]
        p.frameID = frame_id
        sync_p.signalValue = sync_c.signalValue = sync_c.waitValue = seq & 0xFFFFFFFF
        st = C.c_void_p(stream)
        pr, cr, fbr = C.byref(p), C.byref(camera), C.byref(self.fb)
        rc = 0
        if seq > 2:  # partial buffer b was last read by the composites of frame seq-2
            rc |= L.dvr_wait_flags(h["resolved"], h["W"], C.c_uint32((seq - 2) & 0xFFFFFFFF), h["err"], st)
        rc |= L.dvr_render_partial_sync(pr, cr, self.instance, h["mine_rg"][b], h["mine_dp"][b], C.byref(sync_p), st)
        rc |= L.dvr_composite_resolve_peers_sync(pr, cr, self.instance, h["rg"][b], h["dp"][b], h["W"], h["obj"], h["inst"],
                                                 fbr, h["lo"], h["hi"], C.byref(sync_c), st)
        if me == 0 and wait_display:  # the display rank's stream continues once every strip has landed
            rc |= L.dvr_wait_flags(h["resolved"], h["W"], C.c_uint32(seq & 0xFFFFFFFF), h["err"], st)
        if rc != 0:
            raise RuntimeError(f"sort-last frame {seq}: {capi.last_error()}")

    def _render_fused(self, frame_id, camera, stream, b, seq, wait_display):
        """One dvr_render_slab_frame launch: march + exchange + composite + resolve + background strip."""
        capi, me, W = self.capi, self.rank, self.world
        if self._hot is None:
            import ctypes as C
            R = self.MAX_REGIONS
            x = capi.DvrSlabExchange()
            x.nRanks, x.rank, x.maxRegions = W, me, R
            keep = {
                "rg": [(C.c_void_p * W)(*self.rgba_ptrs[k]) for k in range(2)],
                "dp": [(C.c_void_p * W)(*self.depth_ptrs[k]) for k in range(2)],
                "rf": (C.c_void_p * W)(*[self.flag_ptrs[r] for r in range(W)]),
                "sf": (C.c_void_p * W)(*[self.flag_ptrs[r] + R * 16 * 4 for r in range(W)]),
            }
            x.regionFlags = C.cast(keep["rf"], C.c_void_p)
            x.resolvedFlags = C.cast(keep["sf"], C.c_void_p)
            x.regionDone = self.region_done.data_ptr()
            x.errorFlag = self.flag_ptrs[me] + (R * 16 + 16) * 4
            self._hot = {"p": self.params(0), "x": x, "keep": keep, "C": C, "L": capi.lib,
                         "obj": C.c_uint32(self.obj_id), "inst": C.c_uint32(self.inst_id)}
        h = self._hot
        C, L, p, x = h["C"], h["L"], h["p"], h["x"]
        p.frameID = frame_id
        x.seq = seq & 0xFFFFFFFF
        x.partialRgba = C.cast(h["keep"]["rg"][b], C.c_void_p)
        x.partialDepth = C.cast(h["keep"]["dp"][b], C.c_void_p)
        x.waitAllResolved = 1 if (me == 0 and wait_display) else 0
        x.timing = self.timing_ptr or None
        rc = L.dvr_render_slab_frame(C.byref(p), C.byref(camera), self.instance, h["obj"], h["inst"], C.byref(self.fb),
                                     C.byref(x), C.c_void_p(stream))
        if rc != 0:
            raise RuntimeError(f"sort-last frame {seq}: {capi.last_error()}")

    def calibrate(self, field, ranges, limits, camera, stream: int, rounds: int = 3, frames: int = 8,
                  fused_rounds: int = 3):
        """Feedback load balancing: times this rank's march of its slab alone (dvr_render_partial, CUDA events), gathers
        every rank's time, moves the cuts to equal measured work (rebalance_slab_ranges) and applies this rank's new
        range with dvr_field_set_owned_slices — no voxel moves, the slabs were created with a margin.  `ranges` is the
        current partition (all ranks), `limits` the ranges the slabs were created with.  Call it at set-up and whenever
        the camera has moved far; returns (ranges, times_ms of the last round, history).

        `fused_rounds` more rounds then balance what the frame time really depends on: the march PHASE of every rank
        inside the fused frame (first CTA -> last tile, %globaltimer stamps of dvr_render_slab_frame), which is not
        the march alone — flags, background chunks and composites that run beside it stretch it differently per rank
        (N = 8, equal marches alone: 170 us on the display rank, 141 us on the last one, profiles/r02_sort_last_fused.md).
        Part of that time does not scale with the slab's thickness, so these rounds are damped."""
        torch, dist, capi = self.torch, self.dist, self.capi
        ranges = [tuple(r) for r in ranges]
        history = []
        times = None
        for _ in range(max(rounds, 1)):
            p = self.params(1)
            for i in range(2):
                capi.render_partial(p, camera, self.instance, self.rgba_ptrs[0][self.rank], self.depth_ptrs[0][self.rank], stream)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(frames):
                capi.render_partial(p, camera, self.instance, self.rgba_ptrs[0][self.rank], self.depth_ptrs[0][self.rank], stream)
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / frames], dtype=torch.float64, device=self.device)
            if self.world > 1:
                allt = [torch.zeros_like(t) for _ in range(self.world)]
                dist.all_gather(allt, t, group=self.group)
                times = [float(v.item()) for v in allt]
            else:
                times = [float(t.item())]
            history.append({"ranges": [list(r) for r in ranges], "march_ms": [round(v, 4) for v in times]})
            new = rebalance_slab_ranges(ranges, times, limits)
            if new == ranges:
                break
            field.set_owned_slices(*new[self.rank])
            ranges = new
        if self.world > 1 and self.fused:
            for _ in range(max(fused_rounds, 0)):
                tbuf = torch.zeros((frames, 8), dtype=torch.int64, device=self.device)
                tbuf[:, 0] = torch.iinfo(torch.int64).max
                for i in range(2):
                    self.render(1 + i, camera, stream)
                for i in range(frames):
                    self.timing_ptr = tbuf[i].data_ptr()
                    self.render(3 + i, camera, stream)
                self.timing_ptr = 0
                torch.cuda.synchronize()
                t = ((tbuf[:, 1] - tbuf[:, 0]).double().median() / 1e6).reshape(1)  # ms
                allt = [torch.zeros_like(t) for _ in range(self.world)]
                dist.all_gather(allt, t, group=self.group)
                ftimes = [float(v.item()) for v in allt]
                history.append({"ranges": [list(r) for r in ranges], "fused_march_ms": [round(v, 4) for v in ftimes]})
                if not all(v > 0.0 for v in ftimes):
                    break
                new = rebalance_slab_ranges(ranges, ftimes, limits, damping=0.7)
                if new == ranges:
                    break
                field.set_owned_slices(*new[self.rank])
                ranges = new
        return ranges, times, history

    def check_errors(self):
        """True when a bounded spin gave up (a producer never signalled)."""
        if self.world == 1:
            return False
        import ctypes
        v = ctypes.c_uint32()
        ctypes.CDLL("libcudart.so").cudaMemcpy(ctypes.byref(v), ctypes.c_void_p(self.flag_ptrs[self.rank] + (self.MAX_REGIONS * 16 + 16) * 4),
                                               ctypes.c_size_t(4), ctypes.c_int(2))
        return v.value != 0

    def color_tensor(self):
        import ctypes
        n32 = self.npx * (4 if self.fmt == self.capi.DVR_FORMAT_FLOAT32_VEC4 else 1)
        out = self.torch.empty(n32, dtype=self.torch.int32, device=self.device)
        ctypes.CDLL("libcudart.so").cudaMemcpy(ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(self.color_ptr),
                                               ctypes.c_size_t(n32 * 4), ctypes.c_int(3))
        return out

    def depth_tensor(self):
        """Display rank: the assembled depth channel (copy)."""
        import ctypes
        out = self.torch.empty(self.npx, dtype=self.torch.float32, device=self.device)
        ctypes.CDLL("libcudart.so").cudaMemcpy(ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(self.display_depth_ptr),
                                               ctypes.c_size_t(self.npx * 4), ctypes.c_int(3))
        return out

    def close(self):
        self.torch.cuda.synchronize()
        if self.host_frame:
            self.host_frame.close(self.dist if self.world > 1 else None)
        if self.world > 1:
            self.dist.barrier(group=self.group)
            for p in self._peer_open:
                self.capi.ipc_close(p)
            self.dist.barrier(group=self.group)
            for p in self._mine:
                self.capi.ipc_free(p)
            if self._color_owned:
                self.capi.ipc_free(self._color_owned)
