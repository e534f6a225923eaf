"""Sample-split rendering across the GPUs of one box: one process per GPU, one NCCL reduce.

Every (pixel, sample index) is independent -- the seed is PCG32(sampleIndex) + x + W*y (shader.comp:948-958) -- so
rank g renders a contiguous slice of the sample-index range for all pixels into a private fp32 SUM image
(Renderer.dispatch_sum), the sum images are reduced to rank 0 (the path's only exchange step), and rank 0 applies
sum / total * apertureSize^2 * ISO (Renderer.finalize).  The union of samples equals the single-GPU run.
"""


def sample_slice(rank, world, total_samples, first_sample=0):
    """Contiguous slice [begin, end) of the sample-index range owned by `rank`; slices tile the range exactly."""
    if not (0 <= rank < world) or total_samples < 0:
        raise ValueError('bad rank/world/total')
    begin = first_sample + (total_samples * rank) // world
    end = first_sample + (total_samples * (rank + 1)) // world
    return begin, end


def chunks(begin, end, spf):
    """Dispatch plan for one rank: (first_sample, n_samples) pieces of at most spf samples."""
    out = []
    s = begin
    while s < end:
        n = min(spf, end - s)
        out.append((s, n))
        s += n
    return out


def render_split(renderer, params, total_samples, spf, image, dist=None, first_sample=0):
    """Render total_samples samples per pixel split over the ranks of `dist` (torch.distributed, already initialised;
    None = single process).  `image` is this rank's (H, W, 4) float32 CUDA tensor; on return rank 0's holds the
    finished XYZ image in the reference's units, other ranks' hold their partial sums."""
    rank = dist.get_rank() if dist is not None else 0
    world = dist.get_world_size() if dist is not None else 1
    import torch
    # The renderer launches on its OWN stream (created cudaStreamNonBlocking: it never synchronises implicitly with
    # torch's), while image.zero_() and the NCCL reduce are enqueued on torch's current stream.  Every hand-over between
    # the two is ordered explicitly: zero -> dispatches, dispatches -> reduce, reduce -> finalize.
    image.zero_()
    torch.cuda.current_stream().synchronize()
    renderer.bind_image(image)
    begin, end = sample_slice(rank, world, total_samples, first_sample)
    for s, n in chunks(begin, end, spf):
        renderer.dispatch_sum(params, s, n)
    renderer.sync()
    if dist is not None and world > 1:
        dist.reduce(image, dst=0, op=dist.ReduceOp.SUM)
        torch.cuda.current_stream().synchronize()   # the reduce has landed before finalize touches the image
    if rank == 0:
        renderer.finalize(params, total_samples)
        renderer.sync()
    return image
