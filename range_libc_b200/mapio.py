"""Host-side map ingest mirroring the reference's OMap(filename, threshold)
(/root/reference/includes/RangeLib.h:159-201 + RangeUtils.h:30-32).  PNG decoding itself is not
on the accelerated path; this keeps `PyOMap("map.png")` working."""
import numpy as np


def occupancy_from_rgba(rgba, threshold=128.0):
    """rgba: uint8 [rows, cols, 4] as lodepng_decode32 / PIL 'RGBA' produce it.
    The reference reads r = byte 2, g = byte 1, b = byte 0 (:193-195), computes
    gray = (int)(float)(0.229*r + 0.587*g + 0.114*b) (double arithmetic narrowed to float on return,
    then truncated) and marks the cell occupied iff gray < threshold.
    Returns uint8 [W, H] x-major (x = image column, y = image row)."""
    r = rgba[:, :, 2].astype(np.float64)
    g = rgba[:, :, 1].astype(np.float64)
    b = rgba[:, :, 0].astype(np.float64)
    gray = ((0.229 * r + 0.587 * g) + 0.114 * b).astype(np.float32).astype(np.int32)
    return np.ascontiguousarray((gray < threshold).T, dtype=np.uint8)


def load_png(path, threshold=128.0):
    from PIL import Image
    if isinstance(path, bytes):
        path = path.decode()
    return occupancy_from_rgba(np.asarray(Image.open(path).convert("RGBA"), dtype=np.uint8), threshold)


def save_png(path, occ_xmajor):
    """OMap::save (RangeLib.h:264-291): RGBA8 image [rows = H][cols = W], (0,0,0,255) where grid[x][y] is set and
    (255,255,255,255) elsewhere.  Returns False on success / True on error (the reference returns lodepng's code)."""
    from PIL import Image
    if isinstance(path, bytes):
        path = path.decode()
    occ = np.asarray(occ_xmajor)
    img = np.full((occ.shape[1], occ.shape[0], 4), 255, np.uint8)
    img[occ.T != 0, :3] = 0
    try:
        Image.fromarray(img, "RGBA").save(path, format="PNG")
    except Exception as ex:  # noqa: BLE001
        print("encoder error: %s" % ex)
        return True
    return False
