"""Shape arithmetic of the reference, used by the kernel-level tests to fill the C-ABI descriptors.
(The product's host layer classes do the same in C++; these are the test-side restatement.)"""


def conv_pads(w, h, kw, kh, dw, dh, sw, sh, pl, pr, pt, pb):
    """Convolution::make_padding, src/layer/convolution.cpp:328-372 -> resolved (pl, pr, pt, pb)"""
    kext_w = dw * (kw - 1) + 1
    kext_h = dh * (kh - 1) + 1
    if pl > 0 or pr > 0 or pt > 0 or pb > 0:
        return pl, pr, pt, pb
    if pl == -233 and pr == -233 and pt == -233 and pb == -233:
        wpad = kext_w + (w - 1) // sw * sw - w
        hpad = kext_h + (h - 1) // sh * sh - h
        if wpad > 0 or hpad > 0:
            return wpad // 2, wpad - wpad // 2, hpad // 2, hpad - hpad // 2
        return 0, 0, 0, 0
    if pl == -234 and pr == -234 and pt == -234 and pb == -234:
        wpad = kext_w + (w - 1) // sw * sw - w
        hpad = kext_h + (h - 1) // sh * sh - h
        if wpad > 0 or hpad > 0:
            return wpad - wpad // 2, wpad // 2, hpad - hpad // 2, hpad // 2
        return 0, 0, 0, 0
    return 0, 0, 0, 0


def conv_out(w, h, kw, kh, dw, dh, sw, sh, pads):
    pl, pr, pt, pb = pads
    kext_w = dw * (kw - 1) + 1
    kext_h = dh * (kh - 1) + 1
    return (w + pl + pr - kext_w) // sw + 1, (h + pt + pb - kext_h) // sh + 1


def pool_geometry(w, h, kw, kh, sw, sh, pl, pr, pt, pb, pad_mode):
    """Pooling::make_padding + forward, src/layer/pooling.cpp:188-199, :350-412.
    -> dict(outw, outh, pad_left, pad_top (applied), area = (x0, x1, y0, y1) in input coordinates)"""
    wtail = htail = 0
    if pad_mode == 0:
        wt = (w + pl + pr - kw) % sw
        ht = (h + pt + pb - kh) % sh
        wtail = sw - wt if wt != 0 else 0
        htail = sh - ht if ht != 0 else 0
        al, ar, at, ab = pl, pr + wtail, pt, pb + htail
    elif pad_mode == 1:
        al, ar, at, ab = pl, pr, pt, pb
    else:
        wpad = kw + (w - 1) // sw * sw - w
        hpad = kh + (h - 1) // sh * sh - h
        al = ar = at = ab = 0
        if wpad > 0 or hpad > 0:
            if pad_mode == 2:
                at, ab, al, ar = hpad // 2, hpad - hpad // 2, wpad // 2, wpad - wpad // 2
            else:
                at, ab, al, ar = hpad - hpad // 2, hpad // 2, wpad - wpad // 2, wpad // 2
    bw, bh = w + al + ar, h + at + ab
    outw, outh = (bw - kw) // sw + 1, (bh - kh) // sh + 1
    # avg divisor region (pooling.cpp:283-300): bordered sx in [pad_left, bw - pad_right - wtailpad) with the MEMBER pads
    x0 = pl - al
    x1 = (bw - pr - (wtail if pad_mode == 0 else 0)) - al
    y0 = pt - at
    y1 = (bh - pb - (htail if pad_mode == 0 else 0)) - at
    return dict(outw=outw, outh=outh, pad_left=al, pad_top=at, area=(x0, x1, y0, y1))
