"""Slab-sharded paint -> distributed R2C FFT -> P(k) multipoles (one process per GPU).

SURVEY.md section 8e / BASELINE.json configs[3].  The reference has no multi-device code; this
module is the host side that strings the per-rank C-ABI stages of include/jps.h
("slab-sharded mesh") together with three exchanges done through torch.distributed (NCCL over
NVLink on GPUs; gloo in the CPU choreography test):

  paint   : every rank deposits ITS particles (x in its slab) into [1 + N/P + 2][N][N] planes
            (one ghost plane below, two above: enough for CIC/TSC/PCS stencils anchored at their
            lowest node)                                            -> ring halo exchange + add
  FFT     : 2-D R2C of the owned planes, pack by destination, ONE all-to-all of
            (P-1)/P^2 * 8 N^2 (N/2+1) bytes per rank, 1-D C2C along x -> delta_k stays y-sharded
  binning : local fold/bin kernel                                   -> allreduce of nb*3 float64

The exchange helpers are plain functions on tensors so that the choreography (who sends which
block where) is unit-tested on CPU with gloo, independently of the CUDA kernels.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from ._lib import check, lib
from .correlations import HostPipeline, PaintPowspec, _edge_ptr, _host_edges, host_chunks
from .mas import new_paint_workspace, paint_workspace_bytes
from .plan import check_particles, ptr, stream_ptr

GHOST_LO, GHOST_HI = 1, 2


# --------------------------------------------------------------------------- exchanges
def _world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def halo_exchange_add(mesh: torch.Tensor, nxl: int) -> None:
    """mesh: [GHOST_LO + nxl + GHOST_HI, N, N].  Ring exchange: the low ghost plane is added to the
    previous rank's last owned plane, the high ghost planes to the next rank's first owned planes."""
    rank, world = _world()
    lo = mesh[0:GHOST_LO]
    hi = mesh[GHOST_LO + nxl: GHOST_LO + nxl + GHOST_HI]
    own_first = mesh[GHOST_LO: GHOST_LO + GHOST_HI]
    own_last = mesh[GHOST_LO + nxl - GHOST_LO: GHOST_LO + nxl]
    if world == 1:
        own_last += lo
        own_first += hi
        return
    prev, nxt = (rank - 1) % world, (rank + 1) % world
    from_next = torch.empty_like(lo)       # next rank's low ghost lands on my last owned plane
    from_prev = torch.empty_like(hi)       # previous rank's high ghosts land on my first owned planes
    ops = [dist.P2POp(dist.isend, lo.contiguous(), prev), dist.P2POp(dist.isend, hi.contiguous(), nxt),
           dist.P2POp(dist.irecv, from_next, nxt), dist.P2POp(dist.irecv, from_prev, prev)]
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    own_last += from_next
    own_first += from_prev


def transpose_all_to_all(send: torch.Tensor, recv: torch.Tensor) -> None:
    """send: [P][nxl][nyl][nz] (block q goes to rank q); recv: [P][nxl][nyl][nz] = [N][nyl][nz] with the
    block of source rank q at x-planes [q*nxl, (q+1)*nxl)."""
    _, world = _world()
    if world == 1:
        recv.copy_(send)
        return
    s = torch.view_as_real(send) if send.is_complex() else send
    r = torch.view_as_real(recv) if recv.is_complex() else recv
    dist.all_to_all_single(r.view(-1), s.reshape(-1))


def pack_blocks_torch(yz: torch.Tensor, world: int) -> torch.Tensor:
    """Reference implementation of jps_slab_pack with torch ops (used by the CPU choreography test):
    [nxl][N][nz] -> [P][nxl][nyl][nz]."""
    nxl, n, nz = yz.shape
    nyl = n // world
    return yz.reshape(nxl, world, nyl, nz).permute(1, 0, 2, 3).contiguous()


def slab_owner(x: torch.Tensor, xmin: float, box_size: float, n_mesh: int, world: int) -> torch.Tensor:
    """Rank that must paint each particle: owner of the x-plane floor(pos), pos = (x - xmin) * inv with
    inv = 1 / (box_size / n_mesh) evaluated in float32 exactly as the painters do
    (csrc/paint_common.cuh ``grid_pos``, /root/reference/src/mas.py:100-105) -- a particle a rounding
    away from a slab boundary must land on the rank whose ghost planes cover its stencil."""
    inv = np.float32(1.0) / (np.float32(box_size) / np.float32(n_mesh))
    pos = (x - float(np.float32(xmin))) * float(inv)
    cell = torch.floor(pos).to(torch.int64).remainder_(n_mesh)
    return torch.div(cell, n_mesh // world, rounding_mode="floor")


def route_particles(x, y, z, w, box_size: float, n_mesh: int, xmin: float = 0.0):
    """Send every particle to the rank that owns its x-plane (variable-size all-to-all).
    Catalogues generated in place (BASELINE configs[3]) skip this."""
    rank, world = _world()
    if world == 1:
        return x, y, z, w
    owner = slab_owner(x, xmin, box_size, n_mesh, world)
    order = torch.argsort(owner)
    counts = torch.bincount(owner, minlength=world)
    recv_counts = torch.empty_like(counts)
    dist.all_to_all_single(recv_counts, counts)
    s_split, r_split = counts.tolist(), recv_counts.tolist()
    out = []
    for t in (x, y, z, w):
        if t is None:
            out.append(None)
            continue
        src = t[order].contiguous()
        dst = torch.empty(int(sum(r_split)), dtype=t.dtype, device=t.device)
        dist.all_to_all_single(dst, src, output_split_sizes=r_split, input_split_sizes=s_split)
        out.append(dst)
    return tuple(out)


# --------------------------------------------------------------------------- pipeline
class SlabPipeline:
    """Per-rank object: owns the slab mesh, the two complex transpose buffers and the slab plan."""

    def __init__(self, n_mesh, box_size, k_edges, *, order=2, compat="fixed", method="auto", wrap=True,
                 shot_noise=0.0, rank=None, world=None, device=None, transport="auto", overlap=True,
                 layout="auto", pipeline=False, fft="auto"):
        """transport: how the transpose crosses GPUs -- "p2p": one fused pack + peer-store kernel
        over NVLink peer memory (receive buffers mapped into every rank with CUDA IPC); "nccl":
        pack kernel + ``all_to_all_single``; "auto": p2p when the mapping succeeds, else nccl.
        overlap (p2p only): split the owned planes into chunks and send chunk c on a side stream
        while the 2-D FFT of chunk c+1 runs.
        layout: "xfast" = the peer-store kernel transposes on the way so that the shard arrives as
        [y_local][kz][x] and the 1-D FFT along x is contiguous (p2p only); "xslow" = [x][y_local][kz]
        with a strided FFT; "auto" = xfast with p2p, xslow otherwise.
        fft (x-fast layout only): "pencil" = the owned planes are transformed as C2C of half length along z +
        fused untangle/transpose + contiguous C2C along y; "cufft2d" = cuFFT's batched 2-D R2C plan; "auto" = pencil
        when a rank's planes hold 6 GB or more (the measured crossover at 2048^3).
        pipeline: run deposit, halo exchange, 2-D FFT and peer transfer as one pipeline over pieces of planes
        (p2p + overlap only; see _pipelined_paint_fft).  Off by default: measured on 2 B200s (2048^3) the pipelined
        step takes 83.3 ms against 82.7 ms staged -- the deposit (3 CTAs of 58 KB shared memory per SM) and cuFFT
        cannot share an SM, so overlapping them only interleaves them; results are identical either way."""
        r, w = _world()
        self.rank = r if rank is None else rank
        self.world = w if world is None else world
        self.n, self.box = int(n_mesh), float(box_size)
        if self.n % self.world:
            raise ValueError(f"n_mesh={self.n} must be divisible by the number of ranks {self.world}")
        self.nxl = self.n // self.world
        if self.world > 1 and self.nxl < GHOST_HI:
            raise ValueError("slab thinner than the stencil: use fewer ranks")
        self.nz = self.n // 2 + 1
        self.order, self.compat, self.method, self.wrap = int(order), compat, method, bool(wrap)
        self.shot_noise = float(shot_noise)
        self.edges = _host_edges(k_edges)
        self.nb = self.edges.size - 1
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        d = self.device
        self.single = self.world == 1
        # One real rank (not a virtual rank of the test harness): the whole mesh lives on this GPU, so
        # the step is the single-GPU pipeline -- ONE monolithic 3-D R2C plan instead of 2-D + strided
        # 1-D transforms (2048^3: 93.6 ms of FFT against ~55 ms) and the 8-fold binning kernel.
        self.local = None
        if self.single and rank is None:
            try:
                self.local = PaintPowspec(self.n, self.box, self.edges, order=self.order, compat=self.compat,
                                          method=self.method, shot_noise=self.shot_noise, wrap=self.wrap,
                                          device=self.device)
            except (_lib.JpsError, torch.OutOfMemoryError) as e:     # 3-D plan does not fit: 2-D + 1-D transforms below
                self.local, self._local_error = None, repr(e)
                torch.cuda.empty_cache()
        if self.local is not None:
            self.transport, self.xfast, self.handle = "local", False, None
            self.mesh, self.k3d, self.pk, self.nm = self.local.mesh, self.local.k3d, self.local.pk, self.local.nm
            self.gl = self.gh = 0
            self.nxa, self.x0 = self.n, 0
            self.peer_ptrs, self._ipc_bases = None, []
            return
        self.gl, self.gh = (0, 0) if self.single else (GHOST_LO, GHOST_HI)
        self.nxa = self.nxl + self.gl + self.gh
        self.x0 = self.rank * self.nxl - self.gl
        self.mesh = torch.empty((self.nxa, self.n, self.n), dtype=torch.float32, device=d)
        cshape = (self.world, self.nxl, self.nxl, self.nz)
        self.buf_a = torch.empty(cshape, dtype=torch.complex64, device=d)     # yz-transformed / receive
        # packed send buffer; with one rank pack + all-to-all are the identity and are skipped
        self.buf_b = None if self.single else torch.empty(cshape, dtype=torch.complex64, device=d)
        nbytes = C.c_size_t(0)
        with torch.cuda.device(d):
            check(lib.jps_slab_plan_workspace_bytes(self.n, self.world, C.byref(nbytes)))
            self.ws = torch.empty(nbytes.value + 256, dtype=torch.uint8, device=d)
            base = (self.ws.data_ptr() + 255) // 256 * 256
            h = C.c_void_p(0)
            check(lib.jps_slab_plan_create(self.n, self.world, self.rank, C.c_void_p(base),
                                           C.c_size_t(nbytes.value), C.byref(h)), "jps_slab_plan_create")
        self.handle = h
        self.sums = torch.zeros((self.nb, 3), dtype=torch.float64, device=d)
        self.counts = torch.zeros(self.nb, dtype=torch.int64, device=d)
        self.dc = torch.zeros(1, dtype=torch.float32, device=d)
        self.k3d = torch.empty(self.nb, dtype=torch.float32, device=d)
        self.pk = torch.empty((self.nb, 3), dtype=torch.float32, device=d)
        self.nm = torch.empty(self.nb, dtype=torch.float32, device=d)
        self.pws, self.pws_bytes = None, 0
        self.transport = "nccl"
        self.peer_ptrs, self._ipc_bases = None, []
        if not self.single and transport in ("auto", "p2p") and rank is None and dist.is_initialized():
            ok = self._map_peer_buffers()
            flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=d)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)              # every rank must succeed
            if int(flag.item()) == 1:
                self.transport = "p2p"
            elif transport == "p2p":
                raise _lib.JpsError("SlabPipeline: transport='p2p' requested but peer mapping failed")
        self._sync_flag = torch.zeros(1, dtype=torch.int32, device=d)
        self.overlap = bool(overlap)
        self.chunk_planes = int(lib.jps_slab_chunk_planes(self.handle))
        self._side, self._events = None, None
        self._halo_stream, self._halo_ev = None, None
        self._dep_stream = None
        if fft not in ("auto", "pencil", "cufft2d"):
            raise ValueError("fft must be 'auto', 'pencil' or 'cufft2d'")
        self.fft = fft
        self._zero_stream, self._zero_ev = None, None
        self.pipeline = bool(pipeline)
        self._local_peers = False
        self._force_chunks = False                      # tests: take the chunked path on small meshes
        if layout not in ("auto", "xfast", "xslow"):
            raise ValueError("layout must be 'auto', 'xfast' or 'xslow'")
        if layout == "xfast" and self.transport != "p2p" and rank is None:
            raise _lib.JpsError("SlabPipeline: layout='xfast' needs the p2p transport")
        self._want_xfast = layout in ("auto", "xfast")
        self.xfast = False
        if self.transport == "p2p":
            self._set_layout(self._want_xfast)

    def _set_layout(self, xfast):
        mode = 1 if xfast else 0
        if xfast and self.fft == "pencil" and self.n % 2 == 0:
            mode |= 2
        elif xfast and self.fft == "cufft2d":
            mode |= 4
        check(lib.jps_slab_set_layout(self.handle, mode), "jps_slab_set_layout")
        self.xfast = bool(xfast)

    def use_local_peers(self, pipes, xfast=True):
        """Test harness: all virtual ranks live on THIS device, so their receive buffers are plain
        device pointers and the fused peer-store kernels (both layouts) run without IPC or NCCL."""
        self.peer_ptrs = (C.c_void_p * self.world)(*[q.buf_a.data_ptr() for q in pipes])
        self.transport = "p2p"
        self._local_peers = True
        self._set_layout(xfast)

    def _map_peer_buffers(self) -> bool:
        """Map every rank's receive buffer (buf_a) into this process through CUDA IPC, opened on THIS
        rank's device so that its kernels can store into the peers over NVLink; fills self.peer_ptrs
        (ctypes array of device pointers, indexed by rank)."""
        try:
            # (device, 64-byte cudaIpcMemHandle_t of the allocator block, size, offset of the storage, ...)
            shared = self.buf_a.untyped_storage()._share_cuda_()
            byte_off = int(shared[3]) + self.buf_a.storage_offset() * self.buf_a.element_size()
            raw = bytes(shared[1])
            if len(raw) > 64:                        # newer torch: [version byte][type byte]handle; b'c' = cudaMalloc block
                kind = raw[len(raw) - 65: len(raw) - 64]
                if kind != b"c":
                    raise RuntimeError(f"receive buffer is not a plain cudaMalloc block (type {kind!r}): no cudaIpcMemHandle")
                raw = raw[-64:]
            if len(raw) != 64:
                raise RuntimeError(f"unexpected IPC handle size {len(raw)}")
            handles = [None] * self.world
            dist.all_gather_object(handles, (raw, byte_off))
            ptrs, self._ipc_bases = [], []
            with torch.cuda.device(self.device):
                for q, (h, off) in enumerate(handles):
                    if q == self.rank:
                        ptrs.append(self.buf_a.data_ptr())
                        continue
                    base = C.c_void_p(0)
                    check(lib.jps_ipc_open(h, C.byref(base)), "jps_ipc_open")
                    self._ipc_bases.append(base)
                    ptrs.append(base.value + off)
            self.peer_ptrs = (C.c_void_p * self.world)(*ptrs)
            return True
        except Exception as e:                                       # private torch API / no peer access: use NCCL
            self._p2p_error = repr(e)
            return False

    def close(self):
        for base in getattr(self, "_ipc_bases", []):
            lib.jps_ipc_close(base)
        self._ipc_bases = []
        if getattr(self, "handle", None):
            lib.jps_slab_plan_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- stages (each enqueues on the current stream; exchanges are separate so that a test can
    #      drive several virtual ranks on one device)
    def _paint_call(self, x, y, z, w, xmin, ymin, zmin, phase=_lib.PAINT_PHASE_ALL, tx_begin=0, tx_end=0, method=None):
        x, y, z, w, stride = check_particles(x, y, z, w, self.device)
        npart = x.numel()
        meth = _lib.METHOD[self.method if method is None else method]
        if paint_workspace_bytes(self.n, npart, self.order, meth) > self.pws_bytes:
            self.pws, self.pws_bytes = new_paint_workspace(self.n, npart, self.order, meth, self.device)
        check(lib.jps_paint_slab_phase(self.n, self.x0, self.nxa, ptr(x), ptr(y), ptr(z), ptr(w), stride, npart,
                                       float(xmin), float(ymin), float(zmin), self.box, self.order, int(self.wrap),
                                       _lib.COMPAT[self.compat], _lib.VARIANT_VEC, meth, ptr(self.mesh), ptr(self.pws),
                                       self.pws_bytes, int(phase), int(tx_begin), int(tx_end), stream_ptr()),
              "jps_paint_slab")

    def stage_paint(self, x, y, z, w=None, xmin=0.0, ymin=0.0, zmin=0.0, zero=True):
        """Deposit this rank's particles into its planes (+ ghosts).  zero=False accumulates on top of the
        planes as they are (a catalogue streamed in pieces, SlabHostPipeline)."""
        if zero and self.n >= 512 and x.numel() >= (1 << 18) and self.method != "atomic":
            # zero the planes on a side stream while the bucketing passes (which never touch the mesh) run
            main = torch.cuda.current_stream(self.device)
            if self._zero_stream is None:
                self._zero_stream = torch.cuda.Stream(self.device)
                self._zero_ev = (torch.cuda.Event(), torch.cuda.Event())
            self._zero_ev[0].record(main)
            self._zero_stream.wait_event(self._zero_ev[0])
            with torch.cuda.stream(self._zero_stream):
                self.mesh.zero_()
                self._zero_ev[1].record(self._zero_stream)
            self._paint_call(x, y, z, w, xmin, ymin, zmin, phase=_lib.PAINT_PHASE_BUCKET, method="sorted")
            main.wait_event(self._zero_ev[1])
            self._paint_call(x, y, z, w, xmin, ymin, zmin, phase=_lib.PAINT_PHASE_DEPOSIT, tx_begin=0,
                             tx_end=int(lib.jps_paint_tile_rows(self.nxa)), method="sorted")
            return
        if zero:
            self.mesh.zero_()
        self._paint_call(x, y, z, w, xmin, ymin, zmin)

    def owned(self):
        return self.mesh[self.gl: self.gl + self.nxl]

    def stage_fft_yz_pack(self):
        if self.transport == "p2p":
            return self.stage_fft_yz_p2p()
        check(lib.jps_slab_fft_yz(self.handle, ptr(self.owned()), ptr(self.buf_a), stream_ptr()), "jps_slab_fft_yz")
        if not self.single:
            check(lib.jps_slab_pack(self.handle, ptr(self.buf_a), ptr(self.buf_b), stream_ptr()), "jps_slab_pack")

    def stage_fft_yz_only(self):
        """The batched 2-D R2C of the owned planes alone (bench.py: what the transfer has to hide behind)."""
        out = self.buf_a if self.buf_b is None else self.buf_b
        check(lib.jps_slab_fft_yz(self.handle, ptr(self.owned()), ptr(out), stream_ptr()), "jps_slab_fft_yz")

    def _chunked(self):
        """Planes per piece of the chunked FFT / transfer overlap, or 0.  Chunking pays when a piece is tens
        of MB or more (2048^3 on 2-8 GPUs); on a 512^3 mesh the extra launches and stream hops cost more than
        they hide."""
        if self.transport != "p2p":
            return 0
        big = self._force_chunks or self.buf_b.numel() * 8 >= (1 << 30)
        cp = self.chunk_planes if (self.overlap and big) else 0
        return cp if (cp and cp < self.nxl) else 0

    def stage_fft_yz_p2p(self, halo_done=None):
        """2-D FFT into the local buffer, then ONE kernel that writes every destination's block straight
        into that rank's receive buffer over NVLink.  Ordering: the peers finished reading their
        receive buffers before the previous step's allreduce completed (stream order), and the tiny
        allreduce below makes every rank's stores visible before anyone starts the 1-D FFT.
        halo_done: event after which the first and the last piece of planes are final (the ring halo exchange
        running on its own stream only touches the first two and the last owned plane); the pieces in between
        are transformed and sent first, so the exchange hides behind them."""
        cp = self._chunked()
        if cp:
            # chunked: the transfer of chunk c (side stream) runs under the 2-D FFT of the next chunk.  The side stream
            # has HIGH priority: the peer-store kernel is NVLink bound and needs two CTAs per SM, but at default
            # priority its CTAs queue behind the next piece's transform, which was enqueued first and fills every SM
            # (measured on 8 GPUs: fused stage 9.45 ms = transform 4.11 + transfer 5.46, nothing hidden).
            main = torch.cuda.current_stream(self.device)
            nchunk = self.nxl // cp
            if self._side is None:
                self._side = torch.cuda.Stream(self.device, priority=-1)
            if self._events is None or len(self._events) != nchunk + 1:
                self._events = [torch.cuda.Event() for _ in range(nchunk + 1)]
            order = list(range(1, nchunk - 1)) + [0] + ([nchunk - 1] if nchunk > 1 else [])
            for c in order:
                if halo_done is not None and c == 0:
                    main.wait_event(halo_done)
                check(lib.jps_slab_fft_yz_planes(self.handle, ptr(self.owned()), ptr(self.buf_b), c * cp, cp,
                                                 stream_ptr()), "jps_slab_fft_yz_planes")
                self._events[c].record(main)
                self._side.wait_event(self._events[c])
                with torch.cuda.stream(self._side):
                    check(lib.jps_slab_pack_p2p_planes(self.handle, ptr(self.buf_b), self.peer_ptrs, c * cp, cp,
                                                       stream_ptr()), "jps_slab_pack_p2p_planes")
            self._events[-1].record(self._side)
            main.wait_event(self._events[-1])
        else:
            check(lib.jps_slab_fft_yz(self.handle, ptr(self.owned()), ptr(self.buf_b), stream_ptr()), "jps_slab_fft_yz")
            check(lib.jps_slab_pack_p2p(self.handle, ptr(self.buf_b), self.peer_ptrs, stream_ptr()), "jps_slab_pack_p2p")
        if not self._local_peers:
            dist.all_reduce(self._sync_flag)

    def stage_fft_x(self):
        check(lib.jps_slab_fft_x(self.handle, ptr(self.buf_a), stream_ptr()), "jps_slab_fft_x")

    def local_dc(self):
        """Re rho_hat(0): element (ix=0, yl=0, kz=0) of rank 0's shard."""
        return self.buf_a.view(-1)[0:1].real.to(torch.float32)

    def stage_partial(self, normalise=True):
        check(lib.jps_slab_powspec_partial(self.handle, ptr(self.buf_a), ptr(self.dc), int(bool(normalise)), self.box,
                                           _edge_ptr(self.edges), self.nb, self.order, ptr(self.sums), ptr(self.counts),
                                           stream_ptr()), "jps_slab_powspec_partial")

    def stage_finalize(self):
        check(lib.jps_slab_powspec_finalize(self.handle, self.box, _edge_ptr(self.edges), self.nb, ptr(self.sums),
                                            ptr(self.counts), self.shot_noise, ptr(self.k3d), ptr(self.pk), ptr(self.nm),
                                            stream_ptr()), "jps_slab_powspec_finalize")
        return self.k3d, self.pk, self.nm

    # ---- the distributed call
    def __call__(self, x, y, z, w=None, xmin=0.0, ymin=0.0, zmin=0.0):
        """x, y, z[, w]: this rank's particles (x inside its slab; use route_particles otherwise)."""
        if self.local is not None:
            return self.local(x, y, z, w, xmin, ymin, zmin)
        if self.pipeline and self._can_pipeline(x.numel()):
            self._pipelined_paint_fft(x, y, z, w, xmin, ymin, zmin)
            return self._after_transpose()
        self.stage_paint(x, y, z, w, xmin, ymin, zmin)
        return self.finish()

    def _can_pipeline(self, npart):
        """The deposit / transform / transfer pipeline needs the chunked p2p path, the bucketed painter and
        at least three pieces of planes."""
        cp = self._chunked()
        if not cp or self.nxl // cp < 3 or self._local_peers or self.method == "atomic":
            return False
        return npart >= (1 << 18)                 # what jps_paint's "auto" would bucket anyway

    def _rows_touching(self, p_lo, p_hi):
        """Tile rows (16 allocated planes each + order-1 halo planes above) that write any plane in [p_lo, p_hi]."""
        rows = int(lib.jps_paint_tile_rows(self.nxa))
        first = max(0, (p_lo - (self.order - 1)) // 16)
        last = min(rows - 1, p_hi // 16)
        return range(first, last + 1)

    def _pipelined_paint_fft(self, x, y, z, w, xmin, ymin, zmin):
        """Deposit, ring halo exchange, 2-D FFT and peer transfer as ONE pipeline over pieces of planes.
        The bucketed tiles are ordered along x, so the deposit is issued tile row by tile row on its own stream;
        a piece of planes is transformed (main stream) as soon as the tile rows that write it are done, and sent
        (transfer stream) as soon as it is transformed.  The halo exchange (its own stream) starts after the few
        tile rows that write the ghost planes and the boundary planes, which are deposited FIRST; the first and
        the last piece wait for it and go last.  The deposit (shared-memory bound), cuFFT (HBM bound) and the
        NVLink stores then overlap instead of running back to back."""
        main = torch.cuda.current_stream(self.device)
        cp = self._chunked()
        nchunk = self.nxl // cp
        if self._dep_stream is None:
            self._dep_stream = torch.cuda.Stream(self.device)
            self._halo_stream = self._halo_stream or torch.cuda.Stream(self.device)
            self._side = self._side or torch.cuda.Stream(self.device, priority=-1)
        ev = lambda: torch.cuda.Event()
        self.mesh.zero_()
        self._paint_call(x, y, z, w, xmin, ymin, zmin, phase=_lib.PAINT_PHASE_BUCKET, method="sorted")
        e_bucket = ev(); e_bucket.record(main)
        self._dep_stream.wait_event(e_bucket)
        done = set()

        def deposit(rows):
            """Deposit the not-yet-deposited tile rows of `rows` (dep stream); returns an event after them."""
            todo = sorted(r for r in rows if r not in done)
            with torch.cuda.stream(self._dep_stream):
                i = 0
                while i < len(todo):
                    j = i
                    while j + 1 < len(todo) and todo[j + 1] == todo[j] + 1:
                        j += 1
                    self._paint_call(x, y, z, w, xmin, ymin, zmin, phase=_lib.PAINT_PHASE_DEPOSIT,
                                     tx_begin=todo[i], tx_end=todo[j] + 1, method="sorted")
                    i = j + 1
                e = ev(); e.record(self._dep_stream)
            done.update(todo)
            return e

        # 1. everything the halo exchange reads or adds into: ghost planes and the boundary owned planes
        halo_rows = set(self._rows_touching(0, self.gl + GHOST_HI - 1)) | \
            set(self._rows_touching(self.gl + self.nxl - GHOST_LO, self.nxa - 1))
        e_h = deposit(halo_rows)
        self._halo_stream.wait_event(e_h)
        with torch.cuda.stream(self._halo_stream):
            halo_exchange_add(self.mesh, self.nxl)
            halo_done = ev(); halo_done.record(self._halo_stream)
        # 2. middle pieces first, then the two that need the neighbours' ghost planes
        order = list(range(1, nchunk - 1)) + [0, nchunk - 1]
        last_pack = None
        for c in order:
            e_d = deposit(self._rows_touching(self.gl + c * cp, self.gl + (c + 1) * cp - 1))
            main.wait_event(e_d)
            if c in (0, nchunk - 1):
                main.wait_event(halo_done)
            check(lib.jps_slab_fft_yz_planes(self.handle, ptr(self.owned()), ptr(self.buf_b), c * cp, cp, stream_ptr()),
                  "jps_slab_fft_yz_planes")
            e_f = ev(); e_f.record(main)
            self._side.wait_event(e_f)
            with torch.cuda.stream(self._side):
                check(lib.jps_slab_pack_p2p_planes(self.handle, ptr(self.buf_b), self.peer_ptrs, c * cp, cp,
                                                   stream_ptr()), "jps_slab_pack_p2p_planes")
                last_pack = ev(); last_pack.record(self._side)
        rows = int(lib.jps_paint_tile_rows(self.nxa))
        e_rest = deposit(range(rows))                 # nothing left by construction; keeps the invariant explicit
        main.wait_event(e_rest)
        main.wait_event(last_pack)
        main.wait_event(halo_done)
        dist.all_reduce(self._sync_flag)              # every rank's stores are visible before anyone reads its shard

    def _after_transpose(self):
        self.stage_fft_x()
        if self.rank == 0:
            self.dc.copy_(self.local_dc())
        if self.world > 1:
            dist.broadcast(self.dc, src=0)
        self.stage_partial(normalise=True)
        if self.world > 1:
            dist.all_reduce(self.sums)
        return self.stage_finalize()

    def finish(self):
        """Everything after the deposit: halo exchange, distributed FFT, binning, allreduce."""
        if self.local is not None:
            return self.local.finish()
        if not self.single and self._chunked() and self.nxl // self._chunked() >= 3 and not self._local_peers:
            # ring halo exchange on its own stream, hidden behind the transform + transfer of the middle pieces
            main = torch.cuda.current_stream(self.device)
            if self._halo_stream is None:
                self._halo_stream = torch.cuda.Stream(self.device)
            if self._halo_ev is None:
                self._halo_ev = (torch.cuda.Event(), torch.cuda.Event())
            self._halo_ev[0].record(main)
            self._halo_stream.wait_event(self._halo_ev[0])
            with torch.cuda.stream(self._halo_stream):
                halo_exchange_add(self.mesh, self.nxl)
                self._halo_ev[1].record(self._halo_stream)
            self.stage_fft_yz_p2p(halo_done=self._halo_ev[1])
        else:
            if not self.single:
                halo_exchange_add(self.mesh, self.nxl)
            self.stage_fft_yz_pack()
        if not self.single and self.transport == "nccl":
            transpose_all_to_all(self.buf_b, self.buf_a)
        return self._after_transpose()

    # ---- NVLink reference rate of the transpose (bench.py reports the fused stage against it)
    def probe_transpose(self, xfast: bool, repeats: int = 3) -> float:
        """Milliseconds of ONE peer-store launch alone (no FFT next to it) moving this rank's
        (P-1)/P of the 2-D-transformed planes to the peers: xfast=False is the straight contiguous
        copy kernel (the achievable peer-copy rate of this box), xfast=True the transposing store the
        pipeline uses.  Collective (every rank must call it); overwrites the receive buffers."""
        if self.transport != "p2p":
            raise _lib.JpsError("probe_transpose needs the p2p transport")
        was = self.xfast
        self._set_layout(xfast)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = float("inf")
        for _ in range(repeats):
            dist.barrier()
            torch.cuda.synchronize(self.device)
            e0.record()
            check(lib.jps_slab_pack_p2p(self.handle, ptr(self.buf_b), self.peer_ptrs, stream_ptr()), "jps_slab_pack_p2p")
            e1.record()
            torch.cuda.synchronize(self.device)
            best = min(best, e0.elapsed_time(e1))
        dist.barrier()
        self._set_layout(was)
        return best


class SlabHostPipeline:
    """End-to-end call for a catalogue that lives in HOST memory on every rank (its slab's particles):
    host->device copy in pieces on a copy stream while the compute stream deposits the previous piece,
    then halo exchange, distributed FFT, binning, allreduce, and the device->host read of
    (k3D, Pk3D, Nmodes3D).  This is what bench.py's ``e2e`` times on the sharded path."""

    def __init__(self, pipe: SlabPipeline, n_part_max: int, weighted: bool = False, n_chunks=None):
        self.pipe = pipe
        if pipe.local is not None:                       # one rank: the single-GPU host pipeline
            self.inner = HostPipeline(pipe.local, n_part_max, weighted=weighted, n_chunks=n_chunks)
            return
        self.inner = None
        d = pipe.device
        self.cap = int(n_part_max)
        self.n_chunks = host_chunks(n_chunks, self.cap, pipe.nxa * pipe.n * pipe.n)
        self.dev = [torch.empty(self.cap, dtype=torch.float32, device=d) for _ in range(4 if weighted else 3)]
        self.k3d = torch.empty(pipe.nb, dtype=torch.float32).pin_memory()
        self.pk = torch.empty((pipe.nb, 3), dtype=torch.float32).pin_memory()
        self.nm = torch.empty(pipe.nb, dtype=torch.float32).pin_memory()
        self.copy_stream = torch.cuda.Stream(device=d)
        self.paint_done = None

    def __call__(self, x, y, z, w=None, xmin=0.0, ymin=0.0, zmin=0.0):
        if self.inner is not None:
            return self.inner(x, y, z, w, xmin, ymin, zmin)
        host = [x, y, z] + ([w] if w is not None else [])
        n = len(x)
        if n > self.cap or len(host) > len(self.dev):
            raise ValueError("SlabHostPipeline: catalogue larger than the buffers it was built for")
        host = [h if isinstance(h, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(h, dtype=np.float32))
                for h in host]
        p = self.pipe
        compute = torch.cuda.current_stream(p.device)
        nchunks = min(self.n_chunks, max(1, n // 65536))
        bounds = [n * c // nchunks for c in range(nchunks + 1)]
        ready = []
        with torch.cuda.stream(self.copy_stream):
            if self.paint_done is not None:
                self.copy_stream.wait_event(self.paint_done)
            for c in range(nchunks):
                lo, hi = bounds[c], bounds[c + 1]
                for h, d in zip(host, self.dev):
                    d[lo:hi].copy_(h[lo:hi], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.copy_stream)
                ready.append(ev)
        for c in range(nchunks):
            lo, hi = bounds[c], bounds[c + 1]
            compute.wait_event(ready[c])
            wd = self.dev[3][lo:hi] if w is not None else None
            p.stage_paint(self.dev[0][lo:hi], self.dev[1][lo:hi], self.dev[2][lo:hi], wd, xmin, ymin, zmin, zero=(c == 0))
        self.paint_done = torch.cuda.Event()
        self.paint_done.record(compute)
        k3d, pk, nm = p.finish()
        self.k3d.copy_(k3d, non_blocking=True)
        self.pk.copy_(pk, non_blocking=True)
        self.nm.copy_(nm, non_blocking=True)
        compute.synchronize()
        return self.k3d.numpy().copy(), self.pk.numpy().copy(), self.nm.numpy().copy()


def run_virtual_ranks(pipes, catalogs, xmin=0.0, p2p=None):
    """Drive P SlabPipeline objects that live on ONE device through the distributed algorithm,
    doing the three exchanges with tensor copies (test harness for the per-rank kernels).
    p2p = "xslow" / "xfast": the transpose goes through the fused peer-store kernel instead (the
    "peers" are the other pipes' receive buffers on the same device), in that layout."""
    P = len(pipes)
    if p2p is not None and P > 1:
        for p in pipes:
            p.use_local_peers(pipes, xfast=(p2p == "xfast"))
    for p, (x, y, z, w) in zip(pipes, catalogs):
        p.stage_paint(x, y, z, w, xmin, xmin, xmin)
    if P > 1:
        los = [p.mesh[0:GHOST_LO].clone() for p in pipes]
        his = [p.mesh[p.gl + p.nxl: p.gl + p.nxl + GHOST_HI].clone() for p in pipes]
        for r, p in enumerate(pipes):
            p.mesh[p.gl + p.nxl - GHOST_LO: p.gl + p.nxl] += los[(r + 1) % P]
            p.mesh[p.gl: p.gl + GHOST_HI] += his[(r - 1) % P]
    for p in pipes:
        p.stage_fft_yz_pack()
    if P > 1 and pipes[0].transport != "p2p":
        for q, dst in enumerate(pipes):                  # all-to-all: block q of rank r -> rank q, slot r
            for r, src in enumerate(pipes):
                dst.buf_a[r].copy_(src.buf_b[q])
    for p in pipes:
        p.stage_fft_x()
    dc = pipes[0].local_dc().clone()
    total = torch.zeros_like(pipes[0].sums)
    for p in pipes:
        p.dc.copy_(dc)
        p.stage_partial(normalise=True)
        total += p.sums
    outs = []
    for p in pipes:
        p.sums.copy_(total)
        outs.append(tuple(t.clone() for t in p.stage_finalize()))
    return outs
