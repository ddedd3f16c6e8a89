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
