"""ORACLE (test infrastructure, not product code) for the detect -> NMS -> crop -> RektNet joint
(BASELINE config 5): a CPU restatement of

  * the confidence filter + xywh->xyxy of CVC-YOLOv3/detect.py:84-90,
  * the greedy top-k NMS of CVC-YOLOv3/utils/nms.py:4-61,
  * the RektNet pre-processing of RektNet/utils.py:73-76 (`cv2.resize` to the network size) and
    RektNet/detect.py:32-34 (HWC BGR u8 -> CHW, / 255.0, float32).

`cv2.resize` is a third-party dependency of the reference (opencv-python, unpinned in
RektNet/requirements.txt); its 8-bit INTER_LINEAR path is restated here from the published algorithm
(modules/imgproc/src/resize.cpp: 11-bit fixed-point coefficients, HResizeLinear / VResizeLinear<uchar>,
and the exact-2x-downscale shortcut to INTER_AREA) and pinned against the cv2 4.13 of this image by
tests/test_detect_oracle.py and the committed tests/golden/detect_golden.pt.

The reference has NO code that joins the two networks (the crop step lives in a separate inference
repository); `crop_rect` below is therefore OUR definition of the joint, built from the box mapping of
detect.py:93-96 (x / ratio - pad).

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this file.
"""
import numpy as np
import torch


# ------------------------------------------------------------------------------------------ NMS
def filter_and_corners(det, conf_thres):
    """detect.py:84-90: rows with conf > thres, boxes (cx,cy,w,h) -> (x1,y1,x2,y2) in fp32.
    Returns (rows kept [n] int64, box_corner [n,4], scores [n])."""
    det = det.float()
    rows = torch.nonzero(det[:, 4] > conf_thres).flatten()
    d = det[rows]
    box = torch.zeros((d.shape[0], 4), dtype=torch.float32)
    xy = d[:, 0:2]
    wh = d[:, 2:4] / 2
    box[:, 0:2] = xy - wh
    box[:, 2:4] = xy + wh
    return rows, box, d[:, 4].clone()


def nms(boxes, scores, overlap=0.5, top_k=200):
    """utils/nms.py:4-61 with the sort made STABLE (the reference's `scores.sort(0)` leaves the order of equal
    scores unspecified; the stable ascending order = "among equal scores the later row is visited first" is the
    rule the CUDA kernel implements).  Returns kept indices (into `boxes`) in visiting order."""
    n = boxes.shape[0]
    if n == 0:
        return torch.zeros(0, dtype=torch.long)
    b = boxes.float().numpy()
    x1, y1, x2, y2 = b[:, 0], b[:, 1], b[:, 2], b[:, 3]
    area = (x2 - x1) * (y2 - y1)  # :23 fp32
    order = torch.sort(scores.float(), stable=True, dim=0)[1].numpy()[-top_k:]  # :24-26
    keep = []
    idx = order
    with np.errstate(invalid="ignore", divide="ignore"):
        while idx.size > 0:
            i = idx[-1]
            keep.append(int(i))
            if idx.size == 1:
                break
            idx = idx[:-1]
            xx1 = np.maximum(x1[idx], x1[i])  # clamp(min=x1[i]) :43-46
            yy1 = np.maximum(y1[idx], y1[i])
            xx2 = np.minimum(x2[idx], x2[i])
            yy2 = np.minimum(y2[idx], y2[i])
            w = np.maximum(xx2 - xx1, np.float32(0))
            h = np.maximum(yy2 - yy1, np.float32(0))
            inter = w * h
            union = (area[idx] - inter) + area[i]  # :57
            iou = inter / union
            idx = idx[iou <= np.float32(overlap)]  # NaN (0/0) is dropped, like IoU.le() :60
    return torch.tensor(keep, dtype=torch.long)


def detect_nms(det, conf_thres, nms_thres, top_k=200):
    """One image: det [rows, 5+C] -> (rows of the kept detections in visiting order, their corner boxes, scores)."""
    rows, box, sc = filter_and_corners(det, conf_thres)
    keep = nms(box, sc, nms_thres, top_k)
    return rows[keep], box[keep], sc[keep]


# ------------------------------------------------------------------------------------------ cv2.resize (8U, INTER_LINEAR)
_COEF_BITS = 11
_ONE = 1 << _COEF_BITS


def _cv_round_short(v):
    # saturate_cast<short>(float): cvRound = round half to even
    return np.clip(np.rint(v), -32768, 32767).astype(np.int32)


def _linear_tab(src, dst, clamp=True):
    """Per destination index: source index s and the two 11-bit weights, as resize.cpp computes them
    (float32 arithmetic for the fraction, double for the scale).  Horizontally (clamp=True) indices outside the row
    are clamped AND their fraction zeroed; vertically (clamp=False) resize.cpp keeps index and fraction and clips
    the two ROW indices when it loads them -- the same row then enters twice with weights (1-f, f), which is not
    the same number in the two-step fixed-point arithmetic of the vertical pass."""
    scale = float(src) / float(dst)
    s = np.zeros(dst, np.int32)
    a = np.zeros((dst, 2), np.int32)
    for d in range(dst):
        f = np.float32((d + 0.5) * scale - 0.5)
        i = int(np.floor(f))
        f = np.float32(f - np.float32(i))
        if clamp and i < 0:
            i, f = 0, np.float32(0)
        if clamp and i >= src - 1:
            i, f = src - 1, np.float32(0)
        s[d] = i
        a[d, 0] = _cv_round_short(np.float32((np.float32(1) - f) * np.float32(_ONE)))
        a[d, 1] = _cv_round_short(np.float32(f * np.float32(_ONE)))
    return s, a


def resize_linear_u8(img, dsize):
    """cv2.resize(img, dsize) for uint8 HxWxC, default INTER_LINEAR.  dsize = (width, height)."""
    img = np.ascontiguousarray(img)
    assert img.dtype == np.uint8 and img.ndim == 3
    H, W, C = img.shape
    dw, dh = int(dsize[0]), int(dsize[1])
    if (W, H) == (dw, dh):
        return img.copy()
    if W == 2 * dw and H == 2 * dh:
        # resize.cpp: INTER_LINEAR with an exact 2x2 decimation runs the fast INTER_AREA kernel
        v = img.astype(np.int32)
        return ((v[0::2, 0::2] + v[0::2, 1::2] + v[1::2, 0::2] + v[1::2, 1::2] + 2) >> 2).astype(np.uint8)
    sx, ax = _linear_tab(W, dw)
    sy, ay = _linear_tab(H, dh, clamp=False)
    v = img.astype(np.int32)
    sx1 = np.minimum(sx + 1, W - 1)
    # HResizeLinear: rows of int32 = S[sx]*a0 + S[sx+1]*a1 (a1 == 0 wherever sx+1 would run off the row)
    hrow = v[:, sx, :] * ax[None, :, 0, None] + v[:, sx1, :] * ax[None, :, 1, None]
    S0 = hrow[np.clip(sy, 0, H - 1)]
    S1 = hrow[np.clip(sy + 1, 0, H - 1)]
    b0 = ay[:, 0][:, None, None]
    b1 = ay[:, 1][:, None, None]
    # VResizeLinear<uchar,int,short,FixedPtCast<int,uchar,22>>: the two-step shift of the SIMD kernel
    out = (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


# ------------------------------------------------------------------------------------------ the joint (ours)
def crop_rect(box, ratio, pad_w, pad_h, W, H):
    """Network-input corner box -> integer crop rectangle [x0,x1) x [y0,y1) of the original W x H frame.
    Mapping of detect.py:93-96 (x / ratio - pad) in fp32, then floor / ceil, clipped so the rectangle is never
    empty."""
    f = np.float32
    x0 = f(f(box[0]) / f(ratio)) - f(pad_w)
    y0 = f(f(box[1]) / f(ratio)) - f(pad_h)
    x1 = f(f(box[2]) / f(ratio)) - f(pad_w)
    y1 = f(f(box[3]) / f(ratio)) - f(pad_h)
    ix0 = int(min(max(np.floor(x0), 0), W - 1))
    iy0 = int(min(max(np.floor(y0), 0), H - 1))
    ix1 = int(min(max(np.ceil(x1), ix0 + 1), W))
    iy1 = int(min(max(np.ceil(y1), iy0 + 1), H))
    return ix0, iy0, ix1, iy1


def prep_crop(frame, rect, size=(80, 80)):
    """RektNet/utils.py:73-76 + detect.py:33-34 on a crop: resize -> CHW -> /255.0 -> float32."""
    x0, y0, x1, y1 = rect
    img = resize_linear_u8(frame[y0:y1, x0:x1], size)
    return (img.transpose((2, 0, 1)) / 255.0).astype(np.float32)


def synth_frames(B, H, W, seed=0):
    """Synthetic BGR u8 frames: smooth gradients + noise (so that interpolation errors are visible)."""
    rng = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:H, 0:W]
    out = np.empty((B, H, W, 3), np.uint8)
    for b in range(B):
        base = np.stack([(xx * (b + 1) + yy) % 256, (xx + 2 * yy * (b + 1)) % 256, (xx * yy // 7) % 256], -1)
        out[b] = ((base + rng.randint(0, 64, size=(H, W, 3))) % 256).astype(np.uint8)
    return out


def synth_detections(B, rows, C, seed=0, hot=24, img=416.0, ties=True):
    """Detections like Darknet.forward's eval output [B, rows, 5+C]: mostly low confidence, `hot` clustered
    high-confidence rows per image (overlapping groups so that NMS has work), a few exactly tied scores."""
    g = torch.Generator().manual_seed(seed)
    det = torch.zeros(B, rows, 5 + C)
    det[..., 0:2] = torch.rand(B, rows, 2, generator=g) * img
    det[..., 2:4] = 8 + torch.rand(B, rows, 2, generator=g) * 60
    det[..., 4] = torch.rand(B, rows, generator=g) * 0.7
    det[..., 5:] = torch.rand(B, rows, C, generator=g)
    for b in range(B):
        n = int(torch.randint(0, hot + 1, (1,), generator=g)) if b else hot
        if n == 0:
            continue
        sel = torch.randperm(rows, generator=g)[:n]
        centres = torch.rand(max(1, n // 3), 2, generator=g) * (img - 80) + 40
        for j, r in enumerate(sel.tolist()):
            c = centres[j % centres.shape[0]]
            det[b, r, 0:2] = c + torch.randn(2, generator=g) * 4
            det[b, r, 2:4] = torch.tensor([24.0, 40.0]) + torch.randn(2, generator=g) * 3
            det[b, r, 4] = 0.8 + 0.2 * torch.rand(1, generator=g).item()
        if ties and n >= 4:  # exact ties
            det[b, sel[1], 4] = det[b, sel[0], 4]
            det[b, sel[3], 4] = det[b, sel[2], 4]
    return det


# ------------------------------------------------------------------------------------------ validation metric
def _bbox_iou_plus1(b1, b2):
    """utils/utils.py:163-193, corner branch (+1 pixel convention), broadcasting [n,1,4] x [1,t,4]."""
    ix1, iy1 = torch.max(b1[..., 0], b2[..., 0]), torch.max(b1[..., 1], b2[..., 1])
    ix2, iy2 = torch.min(b1[..., 2], b2[..., 2]), torch.min(b1[..., 3], b2[..., 3])
    inter = torch.clamp(ix2 - ix1 + 1, min=0) * torch.clamp(iy2 - iy1 + 1, min=0)
    a1 = (b1[..., 2] - b1[..., 0] + 1) * (b1[..., 3] - b1[..., 1] + 1)
    a2 = (b2[..., 2] - b2[..., 0] + 1) * (b2[..., 3] - b2[..., 1] + 1)
    return inter / (a1 + a2 - inter + 1e-12)


def compute_ap(recall, precision):
    """utils/utils.py:91-119."""
    mrec = torch.cat((torch.zeros(1), recall, torch.ones(1)))
    mpre = torch.cat((torch.zeros(1), precision, torch.zeros(1)))
    for i in range(len(mpre) - 1, 0, -1):
        mpre[i - 1] = torch.max(mpre[i - 1], mpre[i])
    i = torch.nonzero(mrec[1:] != mrec[:-1])
    return torch.sum((mrec[i + 1] - mrec[i]) * mpre[i + 1])


def average_precision(tp, conf, n_gt):
    """utils/utils.py:58-89 (the sort made stable: equal confidences keep their order)."""
    i = torch.sort(-conf, stable=True)[1]
    tp, conf = tp[i].float(), conf[i].float()
    fpc = torch.cumsum(1 - tp, dim=0)
    tpc = torch.cumsum(tp, dim=0)
    recall_curve = tpc / (n_gt + 1e-16)
    r = tpc[-1] / (n_gt + 1e-16)
    precision_curve = tpc / (tpc + fpc)
    p = tpc[-1] / (tpc[-1] + fpc[-1])
    return compute_ap(recall_curve, precision_curve), r, p


def image_ap(box_corner, probabilities, labels, width, height, iou_thres):
    """validate.py:95-130 for one image, after NMS: box_corner [n,4] / probabilities [n] in NMS order, labels [T,5]
    normalised with zero padding.  Returns (ap, r, p, correct u8 [n]) or None where the reference skips the image."""
    if box_corner.shape[0] == 0:
        return None
    inds = torch.sort(-probabilities, stable=True)[1]
    box_corner, probabilities = box_corner[inds], probabilities[inds]
    labels = labels[(labels[:, 1:5] <= 0).sum(dim=1) == 0]
    if labels.shape[0] == 0:  # `if [] in ious.data.tolist(): continue`
        return None
    x = labels[:, 1:5]
    tb = torch.zeros(x.shape)
    tb[:, 0] = x[:, 0] - x[:, 2] / 2
    tb[:, 1] = x[:, 1] - x[:, 3] / 2
    tb[:, 2] = x[:, 0] + x[:, 2] / 2
    tb[:, 3] = x[:, 1] + x[:, 3] / 2
    tb[:, (0, 2)] *= width
    tb[:, (1, 3)] *= height
    detected = torch.zeros(tb.shape[0], dtype=torch.uint8)
    correct = torch.zeros(box_corner.shape[0], dtype=torch.uint8)
    ious = _bbox_iou_plus1(box_corner.unsqueeze(1), tb.unsqueeze(0))
    best_is = torch.argmax(ious, dim=1)
    for i in range(ious.shape[0]):
        best_i = best_is[i]
        if ious[i, best_i] > iou_thres and detected[best_i] == 0:
            correct[i] = 1
            detected[best_i] = 1
    ap, r, p = average_precision(correct, probabilities, labels.shape[0])
    return float(ap), float(r), float(p), correct


def synth_labels_for(det, B, T, conf_thres, seed=0, img=416.0):
    """Ground-truth labels that partly coincide with the confident detections of `det` (so that TP, FP and misses all
    occur): per image, jittered copies of some confident boxes + random boxes, zero padded to T rows."""
    g = torch.Generator().manual_seed(seed)
    out = torch.zeros(B, T, 5)
    for b in range(B):
        hot = torch.nonzero(det[b, :, 4] > conf_thres).flatten()
        if b % 5 == 4:
            continue  # an image without labels
        n_copy = min(int(hot.numel()), T // 2)
        pick = hot[torch.randperm(hot.numel(), generator=g)[:n_copy]]
        rows = []
        for r in pick.tolist():
            box = det[b, r, 0:4].clone()
            box[0:2] += torch.randn(2, generator=g) * 2.0
            box[2:4] *= 1.0 + 0.1 * torch.randn(2, generator=g)
            rows.append(box / img)
        for _ in range(int(torch.randint(0, 4, (1,), generator=g))):
            rows.append(torch.cat([0.1 + 0.8 * torch.rand(2, generator=g), 0.03 + 0.1 * torch.rand(2, generator=g)]))
        rows = rows[:T]
        for j, bx in enumerate(rows):
            out[b, j, 1:5] = bx.clamp(1e-3, 0.999)
    return out


# ------------------------------------------------------------------------------------------ letterbox front end
# CVC-YOLOv3/detect.py:62-72: calculate_padding (utils/utils.py:36-48) -> torchvision pad(fill=127) -> torchvision
# resize (PIL BILINEAR, i.e. Pillow's two-pass convolution resampler, which widens its support when it shrinks an
# image) -> to_tensor (/255).  Pillow (unpinned in requirements.txt, 12.2 in this image) is third-party: its 8-bit
# resampler is restated here from the published algorithm (src/libImaging/Resample.c: precompute_coeffs,
# normalize_coeffs_8bpc, ImagingResampleHorizontal_8bpc / Vertical_8bpc) and pinned against the installed Pillow /
# torchvision by tests/test_detect_oracle.py and the committed goldens.
PIL_PRECISION_BITS = 32 - 8 - 2


def calculate_padding(orig_height, orig_width, new_height, new_width):
    """utils/utils.py:36-48."""
    if max(orig_height, orig_width) == orig_height:
        new_img_width = orig_height * new_width / new_height
        scale_factor = new_height / orig_height
        pad_h = 0
        pad_w = int((new_img_width - orig_width) / 2)
    else:
        scale_factor = new_width / orig_width
        new_img_height = orig_width * new_height / new_width
        pad_w = 0
        pad_h = int((new_img_height - orig_height) / 2)
    return pad_h, pad_w, scale_factor


def pil_bilinear_coeffs(in_size, out_size):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the BILINEAR filter over the whole axis.
    Returns (xmin int32 [out], xcount int32 [out], coeff int32 [out, ksize])."""
    scale = filterscale = float(in_size) / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 1.0 * filterscale
    ksize = int(np.ceil(support)) * 2 + 1
    xmin_a = np.zeros(out_size, np.int32)
    cnt_a = np.zeros(out_size, np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        k = np.zeros(ksize, np.float64)
        ww = 0.0
        for x in range(xmax):
            a = (x + xmin - center + 0.5) * ss
            if a < 0.0:
                a = -a
            w = 1.0 - a if a < 1.0 else 0.0
            k[x] = w
            ww += w
        for x in range(xmax):
            if ww != 0.0:
                k[x] /= ww
        for x in range(ksize):
            v = k[x] * (1 << PIL_PRECISION_BITS)
            kk[xx, x] = int(-0.5 + v) if k[x] < 0 else int(0.5 + v)
        xmin_a[xx] = xmin
        cnt_a[xx] = xmax
    return xmin_a, cnt_a, kk


def _pil_pass(img, xmin, cnt, kk, axis):
    """One 8-bit resampling pass along `axis` (0 = vertical, 1 = horizontal) of an HxWxC uint8 array."""
    src = np.moveaxis(img.astype(np.int64), axis, 0)
    out = np.empty((len(xmin),) + src.shape[1:], np.int64)
    for o in range(len(xmin)):
        acc = np.full(src.shape[1:], 1 << (PIL_PRECISION_BITS - 1), np.int64)
        for x in range(int(cnt[o])):
            acc += src[xmin[o] + x] * int(kk[o, x])
        out[o] = np.clip(acc >> PIL_PRECISION_BITS, 0, 255)
    return np.moveaxis(out, 0, axis).astype(np.uint8)


def pil_resize_bilinear_u8(img, out_w, out_h):
    """PIL.Image.resize((out_w, out_h), BILINEAR) for an HxWxC uint8 array: horizontal pass, then vertical pass on the
    8-bit intermediate (each pass skipped when the size along it does not change)."""
    H, W, _ = img.shape
    if W != out_w:
        img = _pil_pass(img, *pil_bilinear_coeffs(W, out_w), axis=1)
    if H != out_h:
        img = _pil_pass(img, *pil_bilinear_coeffs(H, out_h), axis=0)
    return img


def letterbox(frame_rgb, new_w, new_h):
    """detect.py:62-72 on one HxWx3 uint8 RGB frame -> (fp32 [3,new_h,new_w] in [0,1], (ratio, pad_w, pad_h))."""
    h, w, _ = frame_rgb.shape
    pad_h, pad_w, ratio = calculate_padding(h, w, new_h, new_w)
    padded = np.full((h + 2 * pad_h, w + 2 * pad_w, 3), 127, np.uint8)
    padded[pad_h:pad_h + h, pad_w:pad_w + w] = frame_rgb
    out = pil_resize_bilinear_u8(padded, new_w, new_h)
    chw = torch.from_numpy(out.transpose(2, 0, 1).copy()).to(torch.float32).div(255)  # to_tensor
    return chw.numpy(), (ratio, pad_w, pad_h)


# ------------------------------------------------------------------------------------------ tile-and-scale input pipeline
# CVC-YOLOv3/utils/datasets.py:143-159 (ts mode): scale_image (utils/utils.py:321-326: PIL resize with ANTIALIAS =
# LANCZOS) -> pre_tile_padding (:376-382) -> torchvision pad(fill=127) -> get_patch_spacings (:384-405) -> get_patch
# (:411-426: PIL crop, which rounds its float box) -> to_tensor; labels: datasets.py:176-186 + utils/utils.py:456-472.
# Pillow's LANCZOS is the same two-pass resampler as BILINEAR with the truncated-sinc filter of support 3
# (Resample.c: lanczos_filter / sinc_filter); restated here and pinned against the installed Pillow by the tests.
import math as _math


def pil_lanczos_coeffs(in_size, out_size):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the LANCZOS filter over the whole axis."""
    scale = filterscale = float(in_size) / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 3.0 * filterscale
    ksize = int(_math.ceil(support)) * 2 + 1
    xmin_a = np.zeros(out_size, np.int32)
    cnt_a = np.zeros(out_size, np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale

    def sinc(x):
        if x == 0.0:
            return 1.0
        x = x * _math.pi
        return _math.sin(x) / x

    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        k = np.zeros(ksize, np.float64)
        ww = 0.0
        for x in range(xmax):
            a = (x + xmin - center + 0.5) * ss
            w = sinc(a) * sinc(a / 3) if -3.0 <= a < 3.0 else 0.0
            k[x] = w
            ww += w
        for x in range(xmax):
            if ww != 0.0:
                k[x] /= ww
        for x in range(ksize):
            v = k[x] * (1 << PIL_PRECISION_BITS)
            kk[xx, x] = int(-0.5 + v) if k[x] < 0 else int(0.5 + v)
        xmin_a[xx] = xmin
        cnt_a[xx] = xmax
    return xmin_a, cnt_a, kk


def pil_resize_lanczos_u8(img, out_w, out_h):
    """PIL.Image.resize((out_w, out_h), LANCZOS) for an HxWxC uint8 array (horizontal pass, then vertical)."""
    H, W, _ = img.shape
    if W != out_w:
        img = _pil_pass(img, *pil_lanczos_coeffs(W, out_w), axis=1)
    if H != out_h:
        img = _pil_pass(img, *pil_lanczos_coeffs(H, out_h), axis=0)
    return img


def pre_tile_padding(img_width, img_height, patch_width, patch_height):
    """utils/utils.py:376-382."""
    vert_pad, horiz_pad = 0, 0
    if img_width < patch_width:
        horiz_pad = _math.ceil((patch_width - img_width) / 2)
    if img_height < patch_height:
        vert_pad = _math.ceil((patch_height - img_height) / 2)
    return vert_pad, horiz_pad


def get_patch_spacings(img_width, img_height, patch_width, patch_height):
    """utils/utils.py:384-405: (patches wide, patches high, total, horizontal offset, vertical offset)."""
    assert img_width >= patch_width and img_height >= patch_height
    hn = _math.ceil(img_width / patch_width)
    h_over = hn * patch_width - img_width
    h_off = 0 if hn == 1 else h_over / (hn - 1)
    vn = _math.ceil(img_height / patch_height)
    v_over = vn * patch_height - img_height
    v_off = 0 if vn == 1 else v_over / (vn - 1)
    return hn, vn, vn * hn, h_off, v_off


def patch_boundary(padded_w, padded_h, patch_width, patch_height, patch_index):
    """utils/utils.py:411-426 (the float boundary get_patch returns; PIL's crop rounds it)."""
    n_wide, _, _, h_off, v_off = get_patch_spacings(padded_w, padded_h, patch_width, patch_height)
    row_position = patch_index % n_wide
    left = patch_width * row_position - h_off * row_position
    col_position = _math.floor(patch_index / n_wide)
    top = patch_height * col_position - v_off * col_position
    return (left, top, left + patch_width, top + patch_height)


def tile_scale(frame_rgb, scale, patch_w, patch_h, patch_index):
    """datasets.py:143-159 on one HxWx3 uint8 RGB frame -> (fp32 [3,patch_h,patch_w] in [0,1], boundary, (horiz_pad,
    vert_pad), n_patches)."""
    h, w, _ = frame_rgb.shape
    new_h, new_w = int(h * scale), int(w * scale)  # scale_image, utils/utils.py:321-326
    scaled = pil_resize_lanczos_u8(frame_rgb, new_w, new_h)
    vert_pad, horiz_pad = pre_tile_padding(new_w, new_h, patch_w, patch_h)
    padded = np.full((new_h + 2 * vert_pad, new_w + 2 * horiz_pad, 3), 127, np.uint8)
    padded[vert_pad:vert_pad + new_h, horiz_pad:horiz_pad + new_w] = scaled
    ph, pw = padded.shape[:2]
    n_patches = get_patch_spacings(pw, ph, patch_w, patch_h)[2]
    boundary = patch_boundary(pw, ph, patch_w, patch_h, patch_index)
    x0, y0 = int(round(boundary[0])), int(round(boundary[1]))  # PIL.Image.crop: map(int, map(round, box))
    x1, y1 = int(round(boundary[2])), int(round(boundary[3]))
    patch = np.zeros((y1 - y0, x1 - x0, 3), np.uint8)  # a crop beyond the image reads zeros (never happens: tiles fit)
    ys, xs = max(y0, 0), max(x0, 0)
    ye, xe = min(y1, ph), min(x1, pw)
    patch[ys - y0:ye - y0, xs - x0:xe - x0] = padded[ys:ye, xs:xe]
    chw = torch.from_numpy(patch.transpose(2, 0, 1).copy()).to(torch.float32).div(255)  # to_tensor
    return chw.numpy(), boundary, (horiz_pad, vert_pad), n_patches


def tile_labels(labels_xyhw, scale, horiz_pad, vert_pad, boundary, patch_w, patch_h, num_targets):
    """Labels of one patch: datasets.py:176-186 (add class column, corner form, scale, pad offset,
    filter_and_offset_labels utils/utils.py:456-472) then :300-304 (centre form, normalised) and the zero padding to
    `num_targets` rows (:313).  labels_xyhw: rows (x, y, h, w) with (x, y) the upper-left corner, as in the CSV."""
    left, top, right, bottom = boundary
    rows = []
    for x, y, hh, ww in labels_xyhw:
        x0, y0, x1, y1 = scale * x + horiz_pad, scale * y + vert_pad, scale * (x + ww) + horiz_pad, scale * (y + hh) + vert_pad
        box_area = float(x1 - x0) * (y1 - y0)
        dx = min(x1, right) - max(x0, left)
        dy = min(y1, bottom) - max(y0, top)
        overlap = float(dx * dy) if (dx >= 0 and dy >= 0) else 0
        if box_area > 0 and (overlap / box_area > 0.5 or overlap > 1000):
            rows.append([0.0, max(x0, left) - left, max(y0, top) - top, min(x1, right) - left, min(y1, bottom) - top])
    out = torch.zeros(num_targets, 5)
    if not rows:
        # filter_and_offset_labels returns zeros((len(labels), 5)): all-zero rows, i.e. padding
        return out
    t = torch.tensor(rows, dtype=torch.float32)
    cx, cy = (t[:, 1] + t[:, 3]) / 2, (t[:, 2] + t[:, 4]) / 2
    bw, bh = t[:, 3] - t[:, 1], t[:, 4] - t[:, 2]
    n = min(len(rows), num_targets)
    out[:n, 1], out[:n, 2], out[:n, 3], out[:n, 4] = cx[:n] / patch_w, cy[:n] / patch_h, bw[:n] / patch_w, bh[:n] / patch_h
    return out
