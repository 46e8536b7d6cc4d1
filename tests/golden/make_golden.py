#!/usr/bin/env python
"""Freeze golden vectors from the UNMODIFIED reference (run in the build container only).

    python tests/golden/make_golden.py

Imports /root/reference/model/{cvig_fov,cvig_baseline}.py through oracle/ref_loader.py,
runs the reference's own hot-path functions on small seeded inputs and stores inputs
and outputs under tests/golden/*.npz.  The reference has no tests or fixtures of its own
(SURVEY.md section 4), so these files are what pins the oracle (tests/test_oracle_golden.py)
and, through it, the CUDA kernels.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref_loader  # noqa: E402

cvig = ref_loader.load("cvig_fov")
base = ref_loader.load("cvig_baseline")
torch.set_num_threads(1)  # deterministic reduction order in the fixtures


def sw_of(fov):
    return int(fov / 360 * 512) // 8


def ref_rank_loop(ov, su):
    """Verbatim control flow of cvig_fov.py:543-552, calling the reference functions."""
    count = su.size(0)
    ranks = np.zeros([count], dtype=int)
    for idx in range(count):
        one = torch.unsqueeze(su[idx, :], 0)
        ori = cvig.correlation(ov, one)
        crop = cvig.crop_overhead(ov, ori, one.shape[3])
        d = torch.squeeze(cvig.l2_distance(crop, one))
        ranks[idx] = torch.sum(torch.le(d, d[idx])).item()
    return ranks


def planted(g, q, fov, seed, noise=0.5, c=16, h=4, w=64):
    gen = torch.Generator().manual_seed(seed)
    sw = sw_of(fov)
    ov = torch.randn(g, c, h, w, generator=gen) * 0.06
    su = torch.randn(q, c, h, sw, generator=gen) * 0.06
    sh = torch.randint(0, w, (q,), generator=gen)
    n = min(g, q)
    cols = (sh[:n].view(n, 1) + torch.arange(sw).view(1, sw)) % w
    su[:n] = torch.gather(ov[:n], 3, cols.view(n, 1, 1, sw).expand(n, c, h, sw)) + noise * su[:n]
    return ov, su, sh


def make_resize():
    """Resize -> ImageNormalization -> PolarTransform of the reference's dataset transform chain (cvig_fov.py:389-392) on
    raw-sized float images, exactly as ImagePairDataset.__getitem__ feeds them (uint8 pixels cast to float32,
    cvig_fov.py:88-91).  With the torchvision of the build container Resize antialiases (its pinned 0.9.1 did not); the
    non-antialiased variant is frozen from the call torchvision 0.9.1 made: F.interpolate(bilinear, align_corners=False)."""
    import torch.nn.functional as F
    gen = torch.Generator().manual_seed(41)
    ov8 = torch.randint(0, 256, (3, 301, 283), generator=gen, dtype=torch.uint8)     # downsample by ~1.1-1.2
    su8 = torch.randint(0, 256, (3, 97, 411), generator=gen, dtype=torch.uint8)      # upsample
    big8 = torch.randint(0, 256, (3, 750, 750), generator=gen, dtype=torch.uint8)    # CVUSA's aerial size, ~2.9x down
    # the 750 x 750 image is not stored (1.7 MB of noise): the tests redraw it from the seed and check this checksum
    blob = {"ov_u8": ov8.numpy(), "su_u8": su8.numpy(), "big_seed": 41, "big_sum": int(big8.long().sum()),
            "big_corner": big8[:, :4, :4].numpy()}
    torch.manual_seed(5)                                                              # the generator Resize draws its start from
    state = torch.get_rng_state()
    start = int(torch.randint(0, 512, ()))
    torch.set_rng_state(state)
    d = cvig.Resize("cvusa", fov=90, random_orientation=True)({"surface": su8.float(), "overhead": ov8.float()})
    blob["pano_start"] = start
    blob["pano_surface"] = d["surface"].numpy()                 # [3,128,128]
    blob["pano_overhead_sub"] = d["overhead"][:, ::3, ::5].numpy()
    d = cvig.PolarTransform()(cvig.ImageNormalization()(d))
    blob["pano_surface_norm_sub"] = d["surface"][:, ::3, ::3].numpy()
    blob["pano_polar_sub"] = d["polar"][:, ::3, ::7].numpy()
    d = cvig.Resize("witw", fov=70)({"surface": su8.float(), "overhead": big8.float()})
    blob["witw_surface_sub"] = d["surface"][:, ::2, ::3].numpy()   # [3,128,99] -> sub
    blob["witw_overhead_sub"] = d["overhead"][:, ::5, ::3].numpy()
    # the pinned torchvision 0.9.1's resize of a float tensor
    blob["noaa_overhead_sub"] = F.interpolate(big8.float()[None], size=(256, 256), mode="bilinear", align_corners=False)[0][:, ::5, ::3].numpy()
    blob["noaa_surface_sub"] = F.interpolate(su8.float()[None], size=(128, 512), mode="bilinear", align_corners=False)[0][:, ::3, ::5].numpy()
    np.savez_compressed(os.path.join(HERE, "resize.npz"), **blob)
    print("resize.npz", os.path.getsize(os.path.join(HERE, "resize.npz")))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "resize":
        return make_resize()
    make_resize()
    # ---- polar transform (a1, a2) ------------------------------------------------
    gen = torch.Generator().manual_seed(7)
    tile = torch.randn(2, 256, 256, generator=gen)
    out = cvig.PolarTransform()({"overhead": tile.clone(), "tag": 1})
    assert set(out.keys()) == {"overhead", "tag", "polar"}
    h_s, w_s, s_o = 128, 512, 256
    xx, yy = np.meshgrid(range(w_s), range(h_s))
    import math
    yy_o = (s_o / 2) + (s_o / 2) * (h_s - 1 - yy) / h_s * np.cos(2 * math.pi * xx / w_s)
    xx_o = (s_o / 2) - (s_o / 2) * (h_s - 1 - yy) / h_s * np.sin(2 * math.pi * xx / w_s)
    # the same grid through the reference's bilinear_interpolate must reproduce 'polar'
    chk = cvig.bilinear_interpolate(tile, xx_o, yy_o)
    assert torch.equal(chk, out["polar"])
    np.savez_compressed(os.path.join(HERE, "polar.npz"), tile=tile.numpy(), polar=out["polar"].numpy(),
                        grid_x_sub=xx_o[::5, ::7], grid_y_sub=yy_o[::5, ::7],
                        grid_x_row0=xx_o[0], grid_y_row0=yy_o[0])
    # generic bilinear_interpolate with out-of-range coordinates (exercises all four clips)
    gen = torch.Generator().manual_seed(8)
    im = torch.randn(3, 20, 31, generator=gen)
    bx = (torch.rand(9, 13, generator=gen, dtype=torch.float64) * 40 - 5).numpy()
    by = (torch.rand(9, 13, generator=gen, dtype=torch.float64) * 30 - 5).numpy()
    bx[0, 0], by[0, 0] = 30.0, 19.0       # exactly on the last pixel: all weights 0
    bx[0, 1], by[0, 1] = 0.0, 0.0
    np.savez_compressed(os.path.join(HERE, "bilinear.npz"), im=im.numpy(), x=bx, y=by,
                        out=cvig.bilinear_interpolate(im, bx, by).numpy())

    # ---- uint8 tile -> ImageNormalization -> PolarTransform (the transform chain upstream of a2) ----
    gen = torch.Generator().manual_seed(9)
    tile8 = torch.randint(0, 256, (3, 256, 256), generator=gen, dtype=torch.uint8)
    tile8[0, 255, 128], tile8[1, 128, 255] = 255, 0            # the taps of the two zero-weight pixels
    d = cvig.ImageNormalization()({"surface": tile8[:, :128, :].clone(), "overhead": tile8.clone()})
    d = cvig.PolarTransform()(d)
    assert d["overhead"].dtype == torch.float32
    np.savez_compressed(os.path.join(HERE, "prep.npz"), tile_u8=tile8.numpy(), norm_rows=d["overhead"][:, ::37, :].numpy(),
                        polar=d["polar"].numpy())

    # ---- correlation / crop / distance (a3-a5) -----------------------------------
    cases = {}
    for name, (g, q, fov, seed) in {
        "fov360": (12, 9, 360, 11), "fov90": (10, 10, 90, 12), "fov70": (7, 5, 70, 13),
        "fov180": (6, 8, 180, 14), "fov6": (5, 4, 6, 15),
    }.items():
        ov, su, sh = planted(g, q, fov, seed)
        cases[name] = (ov, su)
    # argmax ties: a gallery item whose columns are all identical -> every shift ties, index 0 wins
    ov, su, _ = planted(6, 6, 90, 16)
    ov[2] = ov[2][:, :, :1].expand(-1, -1, 64).clone()
    ov[4] = torch.cat((ov[4][:, :, :32], ov[4][:, :, :32]), dim=2)   # period-32 item: shifts s and s+32 tie
    cases["ties"] = (ov, su)
    # zero-norm query and zero-norm gallery item -> NaN distances (no epsilon in the reference)
    ov, su, _ = planted(5, 5, 360, 17)
    su[1] = 0
    ov[3] = 0
    cases["zeronorm"] = (ov, su)
    # semantic-variant shapes are identical after the encoder; a non-default C,H exercises generality
    gen = torch.Generator().manual_seed(18)
    cases["c8h2"] = (torch.randn(6, 8, 2, 64, generator=gen), torch.randn(4, 8, 2, 24, generator=gen))
    blob = {}
    for name, (ov, su) in cases.items():
        ori = cvig.correlation(ov, su)
        crop = cvig.crop_overhead(ov, ori, su.shape[3])
        dist = cvig.l2_distance(crop, su)
        blob[name + "_ov"] = ov.numpy()
        blob[name + "_su"] = su.numpy()
        blob[name + "_ori"] = ori.numpy()
        blob[name + "_dist"] = dist.numpy()
        if name in ("fov90", "ties"):
            blob[name + "_crop"] = crop.contiguous().numpy()
    np.savez_compressed(os.path.join(HERE, "match.npz"), **blob)

    # ---- rank / recall (a6) ---------------------------------------------------------
    blob = {}
    for name, (n, fov, seed, noise) in {"r360": (24, 360, 21, 25.0), "r90": (24, 90, 22, 10.0)}.items():
        ov, su, _ = planted(n, n, fov, seed, noise=noise)
        blob[name + "_ov"] = ov.numpy()
        blob[name + "_su"] = su.numpy()
        blob[name + "_ranks"] = ref_rank_loop(ov, su)
    np.savez_compressed(os.path.join(HERE, "ranks.npz"), **blob)

    # ---- baseline rank loop (a7), verbatim control flow of cvig_baseline.py:453-460 ------
    gen = torch.Generator().manual_seed(31)
    n, d = 48, 1536
    ovb = torch.randn(n, d, generator=gen)
    sub = ovb + 14.0 * torch.randn(n, d, generator=gen)
    ranks = np.zeros([n], dtype=int)
    for idx in range(n):
        one = torch.unsqueeze(sub[idx, :], 0)
        dd = torch.pow(torch.sum(torch.pow(ovb - one, 2), dim=1), 0.5)
        ranks[idx] = torch.sum(torch.le(dd, dd[idx])).item()
    np.savez_compressed(os.path.join(HERE, "baseline.npz"), ov=ovb.numpy().astype(np.float16).astype(np.float32),
                        su=sub.numpy().astype(np.float16).astype(np.float32), ranks_full_precision=ranks)
    # (embeddings are stored rounded to fp16-representable fp32 to keep the file small; ranks are recomputed on those)
    ovb = torch.from_numpy(np.load(os.path.join(HERE, "baseline.npz"))["ov"])
    sub = torch.from_numpy(np.load(os.path.join(HERE, "baseline.npz"))["su"])
    for idx in range(n):
        one = torch.unsqueeze(sub[idx, :], 0)
        dd = torch.pow(torch.sum(torch.pow(ovb - one, 2), dim=1), 0.5)
        ranks[idx] = torch.sum(torch.le(dd, dd[idx])).item()
    np.savez_compressed(os.path.join(HERE, "baseline.npz"), ov=ovb.numpy().astype(np.float16), su=sub.numpy().astype(np.float16),
                        ranks=ranks)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
