"""Multi-GPU plumbing: one process per GPU, frames sharded across ranks.

All-intra frames are independent, so the data path needs no collective: frame i of the job belongs to rank
i mod world.  Inter prediction has exactly one exchange -- every rank that encodes a dependant needs the
reconstructed reference frame (after deblock + SAO) of its producer -- which is a broadcast from the producing
rank.  torch.distributed does the transport (NCCL over NVLink on GPUs, gloo on CPU for the tests).
"""
import torch
import torch.distributed as dist


def shard_frames(num_frames, rank, world):
    """Indices of the frames this rank processes (round-robin, keeps POC order per rank)."""
    return list(range(rank, num_frames, world))


def owner_of(frame_idx, world):
    return frame_idx % world


def broadcast_reference_frame(frame, frame_idx, world=None):
    """Make the reconstruction of `frame_idx` (a uint8/int16 tensor, same shape on every rank) available everywhere.
    The owner passes its reconstruction, the others a buffer to fill.  Returns the tensor."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return frame
    world = world or dist.get_world_size()
    # 16-bit samples travel as bytes: neither NCCL nor gloo has a 16-bit integer type
    wire = frame.view(torch.uint8) if frame.dtype in (torch.int16, torch.uint16) else frame
    dist.broadcast(wire, src=owner_of(frame_idx, world))
    return frame


def gather_result_sizes(nbytes):
    """Every rank's bitstream-side payload size (the host concatenates chunks in POC order)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [int(nbytes)]
    t = torch.tensor([int(nbytes)], dtype=torch.int64, device="cuda" if dist.get_backend() == "nccl" else "cpu")
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [int(x[0]) for x in out]


# ---------------------------------------------------------------------------------------------- tiles
def tile_grid(width, height, cols, rows, ctu=64):
    """Uniform tile boundaries in luma samples, split on CTU columns/rows exactly like the reference
    (encoder.c:383-391: boundary i = (i * size_in_ctus) / count).  Returns (x_edges, y_edges)."""
    wc, hc = (width + ctu - 1) // ctu, (height + ctu - 1) // ctu
    xs = [min(width, (i * wc // cols) * ctu) for i in range(cols)] + [width]
    ys = [min(height, (j * hc // rows) * ctu) for j in range(rows)] + [height]
    return xs, ys


def tile_of_rank(rank, cols, rows):
    """Tile t -> GPU t (raster order), 8 tiles of a 4x2 grid on 8 GPUs."""
    t = rank % (cols * rows)
    return t % cols, t // cols


def allgather_tile_reconstructions(frame, width, height, cols, rows):
    """Every rank has reconstructed its own tile inside `frame` (planar I420 tensor of the whole picture, uint8 or
    int16); after the call every rank holds the complete reconstruction -- the reference waits for ALL tiles of the
    previous frame before it starts a dependant (encoderstate.c:1007-1010).  One all_gather of the padded tile
    payloads; ranks beyond cols*rows contribute nothing."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return frame
    world, rank = dist.get_world_size(), dist.get_rank()
    if world < cols * rows:
        raise ValueError(f"{cols}x{rows} tiles need {cols * rows} ranks (one tile per rank), have {world}: the frame would stay incomplete")
    xs, ys = tile_grid(width, height, cols, rows)
    planes = [(0, width, height, 1), (width * height, width // 2, height // 2, 2), (width * height * 5 // 4, width // 2, height // 2, 2)]

    def tile_views(t, buf):
        tx, ty = t % cols, t // cols
        out = []
        for off, pw, ph, sub in planes:
            x0, x1, y0, y1 = xs[tx] // sub, xs[tx + 1] // sub, ys[ty] // sub, ys[ty + 1] // sub
            out.append(buf[off:off + pw * ph].view(ph, pw)[y0:y1, x0:x1])
        return out

    ntiles = cols * rows
    sizes = [sum(v.numel() for v in tile_views(t, frame)) for t in range(ntiles)]
    cap = max(sizes)
    mine = torch.zeros(cap, dtype=frame.dtype, device=frame.device)
    if rank < ntiles:
        mine[:sizes[rank]] = torch.cat([v.reshape(-1) for v in tile_views(rank, frame)])
    gathered = [torch.empty_like(mine) for _ in range(world)]
    if mine.dtype in (torch.int16, torch.uint16):          # 16-bit samples travel as bytes (no 16-bit integer type in NCCL / gloo)
        dist.all_gather([g.view(torch.uint8) for g in gathered], mine.view(torch.uint8))
    else:
        dist.all_gather(gathered, mine)
    for t in range(min(ntiles, world)):
        if t == rank:
            continue
        pos = 0
        for v in tile_views(t, frame):
            v.copy_(gathered[t][pos:pos + v.numel()].view(v.shape))
            pos += v.numel()
    return frame


# ---------------------------------------------------------------------------------------------- measuring the exchanges
def measure_exchanges(device, iters=5, tile_res=(7680, 4320), frame_res=(3840, 2160)):
    """Times the two exchanges of SURVEY 8e on the current process group (every rank must call it):
      * allgather_tile_reconstructions of one 10-bit I420 picture of `tile_res` (config 5: 7680x4320, 99.5 MB) with one
        tile per rank (cols x rows = the largest grid <= world of 4x2 / 2x2 / 2x1), verified: every rank must end up
        with all tiles;
      * broadcast_reference_frame of one 8-bit I420 picture of `frame_res` from its owner rank.
    Device time (CUDA events; wall clock on CPU/gloo), max over ranks.  Returns a dict (same on every rank)."""
    import time
    world, rank = dist.get_world_size(), dist.get_rank()
    on_gpu = device.type == "cuda"

    def timed(fn):
        fn()                                              # warm-up (NCCL channel set-up)
        dist.barrier()
        if on_gpu:
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / iters
        else:
            t = time.perf_counter()
            for _ in range(iters):
                fn()
            ms = (time.perf_counter() - t) * 1e3 / iters
        t = torch.tensor([ms], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    out = {"backend": dist.get_backend(), "ranks": world}
    # ---- tiles
    cols, rows = (4, 2) if world >= 8 else ((2, 2) if world >= 4 else (2, 1))
    w, h = tile_res
    n = w * h * 3 // 2
    frame = torch.zeros(n, dtype=torch.int16, device=device)
    xs, ys = tile_grid(w, h, cols, rows)

    def fill_own_tile():
        if rank < cols * rows:
            tx, ty = rank % cols, rank // cols
            frame[:w * h].view(h, w)[ys[ty]:ys[ty + 1], xs[tx]:xs[tx + 1]] = rank + 1
            for off in (w * h, w * h * 5 // 4):
                frame[off:off + w * h // 4].view(h // 2, w // 2)[ys[ty] // 2:ys[ty + 1] // 2, xs[tx] // 2:xs[tx + 1] // 2] = rank + 1

    fill_own_tile()
    ms = timed(lambda: allgather_tile_reconstructions(frame, w, h, cols, rows))
    # every sample must now carry (its tile's rank + 1): sum over the luma plane = sum of tile areas * (t + 1)
    want = sum((xs[t % cols + 1] - xs[t % cols]) * (ys[t // cols + 1] - ys[t // cols]) * (t + 1) for t in range(cols * rows))
    ok = int(frame[:w * h].to(torch.int64).sum().item()) == want and int((frame[w * h:] == 0).sum().item()) == 0
    okt = torch.tensor([1 if ok else 0], dtype=torch.int64, device=device)
    dist.all_reduce(okt, op=dist.ReduceOp.MIN)
    recv = n * 2 * (cols * rows - 1) / (cols * rows)       # bytes every rank receives from the others
    out["tile_allgather"] = {"collective": "all_gather", "picture": f"{w}x{h} 10-bit I420", "tiles": f"{cols}x{rows}", "bytes_per_picture": n * 2,
                             "ms": ms, "recv_GBps_per_rank": recv / ms / 1e6, "verified": bool(okt.item())}
    del frame
    # ---- reference-frame broadcast
    fw, fh = frame_res
    m = fw * fh * 3 // 2
    ref = torch.full((m,), rank, dtype=torch.uint8, device=device)
    state = {"i": 0}

    def bc():
        if rank == owner_of(state["i"], world):
            ref.fill_(rank)                               # the owner's "reconstruction": its rank in every sample
        broadcast_reference_frame(ref, state["i"], world)
        state["i"] += 1

    ms = timed(bc)
    last_owner = owner_of(state["i"] - 1, world)
    okb = torch.tensor([1 if int(ref[0].item()) == last_owner and int(ref[-1].item()) == last_owner else 0], dtype=torch.int64, device=device)
    dist.all_reduce(okb, op=dist.ReduceOp.MIN)
    out["reference_broadcast"] = {"collective": "broadcast", "picture": f"{fw}x{fh} 8-bit I420", "bytes_per_picture": m, "ms": ms,
                                  "GBps": m / ms / 1e6, "verified": bool(okb.item())}
    return out
