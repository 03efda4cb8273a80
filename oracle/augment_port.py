"""Test infrastructure: CPU restatement (torch fp32) of the reference's training-time input pipeline,
/root/reference/dataset.py:13-333 (FullDataset, mode="train": ToTensor -> ResizeLongestSideAndPad -> RandomRotate ->
ToGray -> ColorAugmentations -> GaussianBlur -> Normalize), with the random decisions drawn from Python's `random` in
the reference's own order.  Pinned against the unmodified reference classes by tests/test_cpu.py (same seed, same
inputs) wherever /root/reference, torchvision and cv2 are importable; the GPU tests compare
sam2_unet_b200.augment.TrainAugment with it.  Never imported by the product package."""
import random

import torch
import torch.nn.functional as F

MEAN, STD = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]


def draw(size, H, W):
    """dataset.py:52-99 (ResizeLongestSideAndPad), :161-163 (RandomRotate), :191 (ToGray), :217-253
    (ColorAugmentations), :278-281 (GaussianBlur): one dict of decisions, consuming `random` exactly like the
    reference does."""
    p = {}
    if random.random() < 0.5:
        sh, sw = random.uniform(1.0, 1.5), random.uniform(1.0, 1.5)
        pad_h, pad_w = int(round(H * sh)) - H, int(round(W * sw)) - W
        pad_top = random.randint(0, pad_h)
        pad_left = random.randint(0, pad_w)
        # dataset.py:76 hands [left, right, top, bottom] to torchvision's F.pad, which reads [left, top, right, bottom]:
        # the image really gets left = pad_left, top = pad_right, right = pad_top, bottom = pad_bottom
        left, top, right, bottom = pad_left, pad_w - pad_left, pad_top, pad_h - pad_top
        p["geom"] = ("pad", top, left, H + top + bottom, W + left + right)
    else:
        sh, sw = random.uniform(0.5, 1.0), random.uniform(0.5, 1.0)
        nh, nw = max(1, int(round(H * sh))), max(1, int(round(W * sw)))
        y1 = random.randint(0, H - nh)
        x1 = random.randint(0, W - nw)
        p["geom"] = ("crop", y1, x1, nh, nw)
    p["rot"] = 0
    if random.random() < 0.75:
        p["rot"] = random.choice([90, 180, 270]) // 90
    p["gray"] = random.random() < 0.5
    ops = []
    if random.random() < 0.8:
        choice = random.randint(0, 3)
        if choice == 0:
            b, c = random.uniform(0.5, 1.5), random.uniform(0.5, 1.5)
            ops = [("brightness", b), ("contrast", c)]
        elif choice == 1:
            b, c = random.uniform(0.5, 1.5), random.uniform(0.5, 1.5)
            s, h = random.uniform(0.5, 1.5), random.uniform(-0.5, 0.5)
            ops = [("brightness", b), ("contrast", c), ("saturation", s), ("hue", h)]
        elif choice == 2:
            s, h = random.uniform(0.5, 1.5), random.uniform(-0.5, 0.5)
            ops = [("saturation", s), ("hue", h)]
        else:
            ops = [("gamma", random.uniform(0.5, 1.5))]
    p["color"] = ops
    p["blur"] = 0
    if random.random() < 0.2:
        p["blur"] = random.choice([3, 5])
    return p


def _gray(img):
    r, g, b = img.unbind(0)
    return (0.2989 * r + 0.587 * g + 0.114 * b).unsqueeze(0)


def _blend(a, b, ratio):
    return (float(ratio) * a + (1.0 - float(ratio)) * b).clamp(0, 1.0)


def _hue(img, f):
    r, g, b = img.unbind(0)
    maxc, minc = img.max(0).values, img.min(0).values
    eqc = maxc == minc
    cr = maxc - minc
    ones = torch.ones_like(maxc)
    s = cr / torch.where(eqc, ones, maxc)
    div = torch.where(eqc, ones, cr)
    rc, gc, bc = (maxc - r) / div, (maxc - g) / div, (maxc - b) / div
    h = (maxc == r) * (bc - gc) + ((maxc == g) & (maxc != r)) * (2.0 + rc - bc) + ((maxc != g) & (maxc != r)) * (4.0 + gc - rc)
    h = torch.fmod(h / 6.0 + 1.0, 1.0)
    h = (h + f) % 1.0
    v = maxc
    i = torch.floor(h * 6.0)
    fr = h * 6.0 - i
    i = i.to(torch.int32) % 6
    p = (v * (1.0 - s)).clamp(0, 1)
    q = (v * (1.0 - s * fr)).clamp(0, 1)
    t = (v * (1.0 - s * (1.0 - fr))).clamp(0, 1)
    sel = lambda opts: sum((i == k) * o for k, o in enumerate(opts))          # noqa: E731
    return torch.stack((sel((v, q, p, p, t, v)), sel((t, v, v, q, p, p)), sel((p, p, t, v, v, q))), 0)


def gaussian_weights(k):
    sigma = k * 0.15 + 0.35                                     # torchvision F.gaussian_blur with sigma=None
    x = torch.linspace(-(k - 1) * 0.5, (k - 1) * 0.5, steps=k)
    pdf = torch.exp(-0.5 * (x / sigma).pow(2))
    return pdf / pdf.sum()


def apply(p, image_u8, label_u8, size):
    """image_u8 [H, W, 3] uint8 tensor (RGB), label_u8 [H, W] uint8 -> dict(image [3,S,S], label [1,S,S]) fp32."""
    img = image_u8.permute(2, 0, 1).float() / 255.0
    lab = label_u8[None].float() / 255.0
    H, W = lab.shape[-2:]
    kind, a, b, ph, pw = p["geom"]
    if kind == "pad":
        img = F.pad(img, (b, pw - W - b, a, ph - H - a), value=1.0)
        lab = F.pad(lab, (b, pw - W - b, a, ph - H - a), value=0.0)
    else:
        img, lab = img[:, a:a + ph, b:b + pw], lab[:, a:a + ph, b:b + pw]
    scale = size / max(ph, pw)
    nh, nw = int(round(ph * scale)), int(round(pw * scale))
    img = F.interpolate(img[None], size=(nh, nw), mode="bilinear", antialias=True, align_corners=False)[0]
    lab = F.interpolate(lab[None], size=(nh, nw), mode="nearest")[0]
    pt, pl = (size - nh) // 2, (size - nw) // 2
    img = F.pad(img, (pl, size - nw - pl, pt, size - nh - pt))
    lab = F.pad(lab, (pl, size - nw - pl, pt, size - nh - pt))
    if p["rot"]:
        img, lab = torch.rot90(img, p["rot"], (1, 2)), torch.rot90(lab, p["rot"], (1, 2))
    if p["gray"]:
        img = _gray(img).expand(3, -1, -1)
    for name, f in p["color"]:
        if name == "brightness":
            img = _blend(img, torch.zeros_like(img), f)
        elif name == "contrast":
            img = _blend(img, _gray(img).mean(), f)
        elif name == "saturation":
            img = _blend(img, _gray(img), f)
        elif name == "hue":
            img = _hue(img, f)
        else:
            img = (img ** f).clamp(0, 1)
    if p["blur"]:
        k = p["blur"]
        w1 = gaussian_weights(k)
        w2 = (w1[:, None] * w1[None, :]).expand(3, 1, k, k)
        img = F.conv2d(F.pad(img[None], (k // 2,) * 4, mode="reflect"), w2, groups=3)[0]
    img = (img - torch.tensor(MEAN).view(3, 1, 1)) / torch.tensor(STD).view(3, 1, 1)
    return {"image": img.contiguous(), "label": lab.contiguous()}
