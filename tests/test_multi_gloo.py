"""world_size-2 gloo test (CPU) of the multi-GPU host logic: band partition, per-band rendering
exactness (each rank draws only its tile band) and the read-back gather to rank 0."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import port
from skity_b200 import hostlib, multigpu, scene


def test_band_ranges_cover_and_align():
    for h in (1, 16, 17, 600, 4096, 16384):
        for w in (1, 2, 3, 8):
            bands = multigpu.band_ranges(h, w)
            assert bands[0][0] == 0 and bands[-1][1] == h
            for (a0, a1), (b0, b1) in zip(bands, bands[1:]):
                assert a1 == b0 and a0 <= a1
            assert all(y0 % 16 == 0 for y0, _ in bands)
    assert [multigpu.canvas_owner(i, 4) for i in range(6)] == [0, 1, 2, 3, 0, 1]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port_no, dl, h, w, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    bands = multigpu.band_ranges(h, world)
    y0, y1 = bands[rank]
    # each rank rasterises the whole replicated display list but keeps only its band (on the GPU the
    # band is enforced by skb_surface_set_band; here the CPU port stands in for the device)
    full = port.render(dl)
    image = np.zeros_like(full)
    image[y0:y1] = full[y0:y1]
    multigpu.gather_bands(multigpu.host_band_tensor_factory(image), bands, rank, world, dist)
    if rank == 0:
        np.save(out_path, image)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(not os.path.exists(hostlib.LIB_PATH), reason="host plug-in not built")
def test_two_rank_band_gather_reassembles_frame(tmp_path):
    s = scene.scene_random_fills(40, 0, 77, box=120.0, width=200, height=150)
    dl = hostlib.encode_scene(s.encode())
    want = port.render(dl)
    out = str(tmp_path / "gathered.npy")
    mp.spawn(_worker, args=(2, _free_port(), dl, 150, 200, out), nprocs=2, join=True)
    got = np.load(out)
    assert np.array_equal(got, want)


def _worker_shared(rank, world, port_no, dl, h, w, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from skity_b200 import device
    y0, y1 = multigpu.band_ranges(h, world)[rank]
    # every rank gets ITS part of the display list (skb_display_list_cull_rows), renders it (the CPU port stands in for
    # the device) and puts its band into the host image all ranks share; no gather at all
    part = device.cull_display_list_rows(dl, y0, y1)
    shared = multigpu.SharedHostImage(f"skb_test_{port_no}", (h, w, 4), rank, dist)
    shared.rows(y0, y1)[...] = port.render(part)[y0:y1]
    dist.barrier()
    if rank == 0:
        np.save(out_path, np.array(shared.array))
        np.save(out_path + ".sizes.npy", np.array([len(part), len(dl)]))
    dist.barrier()
    shared.close()
    dist.destroy_process_group()


@pytest.mark.skipif(not os.path.exists(hostlib.LIB_PATH), reason="host plug-in not built")
def test_two_rank_culled_lists_and_shared_host_image(tmp_path):
    """The end-to-end multi-GPU path of bench.py on CPU: per-band display lists, every rank writes its band into one
    shared host image (multigpu.SharedHostImage), rank 0 sees the whole frame."""
    s = scene.scene_random_fills(60, 0, 78, box=90.0, width=240, height=320)
    dl = hostlib.encode_scene(s.encode())
    want = port.render(dl)
    out = str(tmp_path / "shared.npy")
    mp.spawn(_worker_shared, args=(2, _free_port(), dl, 320, 240, out), nprocs=2, join=True)
    assert np.array_equal(np.load(out), want)
    part_bytes, all_bytes = np.load(out + ".sizes.npy")
    assert part_bytes < all_bytes
