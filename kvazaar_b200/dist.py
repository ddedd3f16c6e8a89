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
    dist.broadcast(frame, src=owner_of(frame_idx, world))
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
    dist.all_gather(gathered, mine)
    for t in range(min(ntiles, world)):
        if t == rank:
            continue
        pos = 0
        for v in tile_views(t, frame):
            v.copy_(gathered[t][pos:pos + v.numel()].view(v.shape))
            pos += v.numel()
    return frame
