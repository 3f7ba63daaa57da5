"""Multi-GPU partitioning of the raster path (one process per GPU, torch.distributed for plumbing).

The path shards with NO data-path collective:
  * batch of independent canvases  -> canvas i is rendered by rank i % world      (canvas_owner)
  * one large canvas               -> contiguous bands of tile rows, the display list replicated on
                                      every rank, each rank rendering only its band  (band_ranges)
The only collective is the read-back gather of the bands to rank 0 (gather_bands): point-to-point
sends of each band's contiguous rows — NCCL over NVLink on GPUs, gloo on CPU for the tests.
"""
import numpy as np

TILE = 16


def canvas_owner(index, world):
    return index % world


def band_ranges(height, world):
    """Contiguous, tile-aligned row bands [(y0, y1)] * world covering [0, height).  Ranks beyond the
    number of tile rows get empty bands (y0 == y1)."""
    tiles_y = (height + TILE - 1) // TILE
    out = []
    for r in range(world):
        t0 = tiles_y * r // world
        t1 = tiles_y * (r + 1) // world
        out.append((min(t0 * TILE, height), min(t1 * TILE, height)))
    return out


class _CudaBuffer:
    """Exposes raw device memory to torch through __cuda_array_interface__."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def surface_band_tensor(surface, y0, y1):
    """1-D uint8 CUDA tensor aliasing rows [y0, y1) of a device surface (rows are pitch bytes apart,
    so a band is one contiguous block)."""
    import torch
    ptr, pitch = surface.device_ptr()
    n = (y1 - y0) * pitch
    if n == 0:
        return torch.empty(0, dtype=torch.uint8, device="cuda"), pitch
    return torch.as_tensor(_CudaBuffer(ptr + y0 * pitch, n), device="cuda"), pitch


def gather_bands(band_tensor_of, bands, rank, world, dist):
    """Gathers every rank's band into rank 0's full image.

    band_tensor_of(y0, y1) -> 1-D uint8 tensor aliasing rows [y0, y1) of THIS rank's image buffer.
    On rank 0 the other ranks' bands are received straight into its own buffer; other ranks send."""
    if world == 1:
        return
    if rank == 0:
        reqs = []
        for r in range(1, world):
            y0, y1 = bands[r]
            if y1 > y0:
                reqs.append(dist.irecv(band_tensor_of(y0, y1), src=r))
        for q in reqs:
            q.wait()
    else:
        y0, y1 = bands[rank]
        if y1 > y0:
            dist.send(band_tensor_of(y0, y1), dst=0)


def host_band_tensor_factory(image):
    """band_tensor_of for a host (H, W, 4) uint8 numpy image (CPU/gloo path of the tests)."""
    import torch
    flat = torch.from_numpy(image.reshape(-1))
    row = image.shape[1] * image.shape[2]

    def f(y0, y1):
        return flat[y0 * row:y1 * row]
    return f


def fuse_gather_into_fine_pass(surface, rank, dist, root=0):
    """Band split with the gather fused into the fine pass: `root` exports its canvas, every other rank maps it
    (CUDA IPC, peer memory over NVLink) and from then on stores the finished pixels of its band there.  After each
    frame a barrier is all that is left of the gather.  Collective: every rank must call it."""
    box = [surface.export_canvas() if rank == root else None]
    dist.broadcast_object_list(box, src=root)
    if rank != root:
        surface.set_remote_canvas(box[0])


class SharedHostImage:
    """A host image (H, W, 4 uint8) in POSIX shared memory that every rank of one node maps: each rank copies the band
    it rendered straight from its own GPU into its rows (skb_surface_read_pixels_async), so a canvas split over N GPUs
    reaches the host over N PCIe links at once instead of through rank 0's alone — and the whole frame sits in rank
    0's address space.  With `register` the mapping is page-locked for CUDA (cudaHostRegister), which is what makes
    the copies asynchronous and full speed.  Collective: every rank constructs it with the same name."""

    def __init__(self, name, shape, rank, dist=None, register=False, my_rows=None):
        import mmap
        import os
        self.shape = tuple(int(v) for v in shape)
        self.nbytes = int(np.prod(self.shape))
        self.path = os.path.join("/dev/shm", name)
        self.owner = rank == 0
        self._registered = False
        if self.owner:
            fd = os.open(self.path, os.O_CREAT | os.O_RDWR | os.O_TRUNC, 0o600)
            os.ftruncate(fd, self.nbytes)
        if dist is not None:
            dist.barrier()
        if not self.owner:
            fd = os.open(self.path, os.O_RDWR)
        self._mm = mmap.mmap(fd, self.nbytes)
        os.close(fd)
        self.array = np.frombuffer(self._mm, dtype=np.uint8).reshape(self.shape)
        # first touch decides which NUMA node a page lives on: every rank touches the rows it will fill (my_rows), so the
        # DMA writes of a GPU land in memory near the process that drives it instead of all on rank 0's node
        if my_rows is not None:
            self.array[my_rows[0]:my_rows[1]] = 0
        if dist is not None:
            dist.barrier()
        if self.owner and my_rows is None:
            self.array[...] = 0          # touch every page once
        if dist is not None:
            dist.barrier()
        if self.owner:
            os.unlink(self.path)         # the mappings keep it alive
        if register:
            import torch
            rc = torch.cuda.cudart().cudaHostRegister(self.array.ctypes.data, self.nbytes, 0)
            self._registered = int(rc) == 0     # not page-locked: the copies still work, staged by the library

    @property
    def registered(self):
        return self._registered

    @staticmethod
    def room_for(nbytes):
        import shutil
        try:
            return shutil.disk_usage("/dev/shm").free > nbytes + (256 << 20)
        except OSError:
            return False

    def rows(self, y0, y1):
        return self.array[y0:y1]

    def close(self):
        if self._registered:
            import torch
            torch.cuda.cudart().cudaHostUnregister(self.array.ctypes.data)
            self._registered = False
        self.array = None
        try:
            self._mm.close()
        except BufferError:
            pass
