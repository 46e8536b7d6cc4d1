"""The heat-map sweep of the reference (tools/heatmap/heatmap.py:113-187) and the streamed tile preparation behind it.

The reference cuts a satellite strip into tiles with GDAL (out of scope: file I/O), runs every tile through
``ResizeOverhead -> ImageNormalization -> PolarTransform`` on the CPU in a DataLoader worker (heatmap.py:134-145), encodes
the batches, grows ``overhead_embed`` with torch.cat (heatmap.py:161-168) and scores one photo against all tiles
(heatmap.py:171-177).  Here the tiles arrive as batches of raw pixels; every batch is uploaded on a copy stream while the
previous one is being processed, prepared by the fused kernels (uint8 in, normalised polar image out), encoded, and written
once into a preallocated gallery operand; the one-photo sweep is evaluated entirely in fp32 from the gallery's spectra.
"""
import torch

from . import ops


def prefetch_to_device(batches, device=None):
    """Yield the batches of an iterable of host tensors as device tensors; batch i+1 is copied on a copy stream while the
    caller works on batch i.  Pinned host tensors make the copies asynchronous; device tensors pass through."""
    device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    copy_stream = torch.cuda.Stream(device=device)
    it = iter(batches)

    def start(batch):
        if batch.is_cuda:
            return batch, None
        with torch.cuda.stream(copy_stream):
            dev = batch.to(device, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return dev, ev

    try:
        nxt = start(next(it))
    except StopIteration:
        return
    while nxt is not None:
        cur, ev = nxt
        try:
            nxt = start(next(it))
        except StopIteration:
            nxt = None
        if ev is not None:
            torch.cuda.current_stream(device).wait_event(ev)
            cur.record_stream(torch.cuda.current_stream(device))
        yield cur


def prepare_tiles(tiles, mean=ops.IMG_MEAN, std=ops.IMG_STD, divisor=255.0, antialias=True, exact=False):
    """ResizeOverhead -> ImageNormalization -> PolarTransform (heatmap.py:134-138 = cvig_fov.py:133, 147, 186-209) on a device
    batch [n,C,h,w]: uint8 tiles that already have the model's size go through the one-kernel uint8 path (one byte per
    pixel read), anything else through resize + normalise, then the polar kernel.  Returns [n,C,128,512] fp32."""
    size = ops.OVERHEAD_SIZE
    if tiles.dtype == torch.uint8 and tiles.shape[-1] == size and tiles.shape[-2] == size:
        return ops.normalized_polar(tiles, mean, std, divisor, exact=exact)
    norm = ops.resize_normalize(tiles, size, size, antialias, mean, std, divisor)
    return ops.polar_transform(norm, exact=exact)


def streamed_polar(tile_batches, mean=ops.IMG_MEAN, std=ops.IMG_STD, divisor=255.0, device=None, antialias=True):
    """BASELINE configs[4]'s first half: a tile set too large for the device at once (100k 5-channel tiles are 131 GB as
    fp32) streams through the polar transform batch by batch from (pinned) host memory; yields the normalised polar images of
    each batch on the device for the encoder.  cvig_semantic.py:163-176 divides only the image channels by 255: pass
    divisor=(255, 255, 255, 1, 1) with the five-channel mean / std."""
    for batch in prefetch_to_device(tile_batches, device):
        yield prepare_tiles(batch, mean, std, divisor, antialias)


def heatmap_sweep(tile_batches, n_tiles, surface_image, surface_encoder, overhead_encoder, fov=360, output_width_max=64,
                  mean=ops.IMG_MEAN, std=ops.IMG_STD, divisor=255.0, antialias=True, device=None, exact_polar=False):
    """tools/heatmap/heatmap.py:113-187 without the GDAL tiling: one photo against ``n_tiles`` satellite tiles.

    tile_batches: iterable of raw tile batches [n,3,h,w] (uint8 or fp32; host -- preferably pinned -- or device);
    surface_image: the raw photo [3,h,w].  The encoders are the caller's (cvig.FOV_DSM, heatmap.py:148-154).
    Returns (orientation in degrees [n_tiles], dissimilarity [n_tiles], score [n_tiles]) as heatmap.py:171-177 computes them;
    the 1 x G sweep is evaluated in fp32 (csrc/finish.cu: columns_kernel), so they are the reference's own values.
    """
    device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    surface_width = int(fov / 360 * ops.SURFACE_WIDTH_MAX)
    with torch.no_grad(), torch.cuda.device(device):
        photo = surface_image.to(device, non_blocking=True)
        # ResizeSurface + ImageNormalization (heatmap.py:68-101): straight to 128 x surface_width, no column window
        surface = ops.resize_normalize(photo.unsqueeze(0), ops.SURFACE_HEIGHT_MAX, surface_width, antialias, mean, std, divisor)
        surface_embed = surface_encoder(surface)                                   # [1,16,4,sw]
        builder = None
        for polar in streamed_polar(tile_batches, mean, std, divisor, device, antialias) if not exact_polar else (
                prepare_tiles(b, mean, std, divisor, antialias, exact=True) for b in prefetch_to_device(tile_batches, device)):
            part = overhead_encoder(polar)                                         # [n,16,4,64]
            if builder is None:
                builder = ops.GalleryBuilder(n_tiles, surface_embed.shape[3], channels=part.shape[1], height=part.shape[2],
                                             width=part.shape[3], device=device)
            builder.append(part)
        if builder is None or builder.count != n_tiles:
            raise ValueError("heatmap_sweep: expected %d tiles, got %d" % (n_tiles, 0 if builder is None else builder.count))
        gallery = builder.finish()
        queries = ops.QueryBatch(surface_embed, impl=gallery.impl)
        dist = torch.empty((n_tiles, 1), dtype=torch.float32, device=device)
        ori = torch.empty((n_tiles, 1), dtype=torch.int64, device=device)
        ops.exact_columns(gallery, queries, dist=dist, ori64=ori, ld=1)
        orientations = torch.squeeze(ori) * 360 / output_width_max - 180
        distances = torch.squeeze(dist)
        scores = torch.exp(10. * (1. - distances))
    return orientations, distances, scores
