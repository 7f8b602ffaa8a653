"""`coco_gt.json` of the dataset-creation script.

Mirrors scf/segmentation/evaluation/coco_gt.py:14-134 (`COCOGtCreator`: categories, per-class external contours of the
label half of every PNG, one RLE annotation per contour with area and bounding box) and is written by
`create_dataset.main` exactly where the reference writes it (create_dataset_for_segmentation.py:204-206).

The reference delegates polygon -> RLE, area and bbox to pycocotools 2.0.2 (`mask.frPyObjects`, `mask.area`,
`mask.toBbox`; requirements.txt:20), which is in neither this image nor /root/reference.  The three functions are
restated here from the published COCO C API they wrap (cocoapi `common/maskApi.c`: `rleFrPoly`, `rleToString`,
`rleArea`, `rleToBbox`).  PARITY UNPINNED: there is no pycocotools here to generate golden vectors with; the tests pin
the restatement through the API's own identities (a box polygon encodes to area w*h and decodes to that box; string
codec round trip) -- see tests/test_coco_gt.py.
"""
import datetime
from pathlib import Path
from typing import Dict, Iterable, List, Tuple

import cv2
import numpy
from PIL import Image, ImageColor


# ------------------------------------------------------------------------------------------- COCO mask API (RLE)
def rle_from_polygon(xy, h: int, w: int) -> numpy.ndarray:
    """`rleFrPoly` (maskApi.c): run lengths (column-major, starting with a run of zeros) of the polygon x0,y0,x1,y1,...
    Boundary points are traced on a 5x upsampled grid; a pixel is inside iff its centre is."""
    xy = numpy.asarray(xy, dtype=numpy.float64).ravel()
    k = xy.size // 2
    scale = 5.0
    x = (scale * xy[0:2 * k:2] + 0.5).astype(numpy.int64)
    y = (scale * xy[1:2 * k:2] + 0.5).astype(numpy.int64)
    x = numpy.append(x, x[0])
    y = numpy.append(y, y[0])
    us, vs = [], []
    for j in range(k):
        xs, xe, ys, ye = int(x[j]), int(x[j + 1]), int(y[j]), int(y[j + 1])
        dx, dy = abs(xe - xs), abs(ys - ye)
        flip = (dx >= dy and xs > xe) or (dx < dy and ys > ye)
        if flip:
            xs, xe, ys, ye = xe, xs, ye, ys
        if dx >= dy:
            s = (ye - ys) / dx if dx else 0.0
            t = numpy.arange(dx + 1, dtype=numpy.int64)
            if flip:
                t = dx - t
            us.append(t + xs)
            vs.append((ys + s * t + 0.5).astype(numpy.int64))
        else:
            s = (xe - xs) / dy
            t = numpy.arange(dy + 1, dtype=numpy.int64)
            if flip:
                t = dy - t
            vs.append(t + ys)
            us.append((xs + s * t + 0.5).astype(numpy.int64))
    u = numpy.concatenate(us) if us else numpy.zeros(0, dtype=numpy.int64)
    v = numpy.concatenate(vs) if vs else numpy.zeros(0, dtype=numpy.int64)
    # points where the boundary crosses a pixel-column boundary, downsampled
    if u.size > 1:
        u0, u1, v0, v1 = u[:-1], u[1:], v[:-1], v[1:]
        sel = u1 != u0
        xd = numpy.where(u1 < u0, u1, u1 - 1).astype(numpy.float64)
        xd = (xd + 0.5) / scale - 0.5
        sel &= (numpy.floor(xd) == xd) & (xd >= 0) & (xd <= w - 1)
        yd = numpy.minimum(v0, v1).astype(numpy.float64)
        yd = (yd + 0.5) / scale - 0.5
        yd = numpy.ceil(numpy.clip(yd, 0.0, float(h)))
        a = (xd[sel].astype(numpy.int64) * h + yd[sel].astype(numpy.int64))
    else:
        a = numpy.zeros(0, dtype=numpy.int64)
    a = numpy.sort(numpy.append(a, h * w))
    a = numpy.diff(a, prepend=0)
    # zero-length runs cancel: a crossing pair at the same position is no run at all
    b = [int(a[0])]
    j, n = 1, a.size
    while j < n:
        if a[j] > 0:
            b.append(int(a[j]))
            j += 1
        else:
            j += 1
            if j < n:
                b[-1] += int(a[j])
                j += 1
    return numpy.asarray(b, dtype=numpy.uint32)


def rle_to_string(counts) -> bytes:
    """`rleToString`: LEB128-like, 6 bits per character (ASCII 48..111), counts after the third stored as the difference
    to the count two positions back."""
    out = bytearray()
    counts = [int(c) for c in counts]
    for i, c in enumerate(counts):
        x = c - counts[i - 2] if i > 2 else c
        more = True
        while more:
            ch = x & 0x1f
            x >>= 5                                   # arithmetic shift, as on a C long
            more = (x != -1) if (ch & 0x10) else (x != 0)
            if more:
                ch |= 0x20
            out.append(ch + 48)
    return bytes(out)


def rle_from_string(s) -> List[int]:
    """`rleFrString` (inverse of rle_to_string; used by the tests and by consumers of coco_gt.json)."""
    if isinstance(s, str):
        s = s.encode('ascii')
    counts, p = [], 0
    while p < len(s):
        x, k, more = 0, 0, True
        while more:
            c = s[p] - 48
            x |= (c & 0x1f) << (5 * k)
            more = bool(c & 0x20)
            p += 1
            k += 1
            if not more and (c & 0x10):
                x |= -1 << (5 * k)
        if len(counts) > 2:
            x += counts[-2]
        counts.append(x)
    return counts


def rle_area(counts) -> int:
    """`rleArea`: the foreground runs are the odd-indexed counts."""
    return int(sum(int(c) for c in list(counts)[1::2]))


def rle_to_bbox(counts, h: int, w: int) -> List[float]:
    """`rleToBbox`: [x, y, width, height] as doubles."""
    counts = [int(c) for c in counts]
    m = (len(counts) // 2) * 2
    if m == 0:
        return [0.0, 0.0, 0.0, 0.0]
    xs, ys, xe, ye, cc, xp = w, h, 0, 0, 0, 0
    for j in range(m):
        cc += counts[j]
        t = cc - j % 2
        y = t % h
        x = (t - y) // h
        if j % 2 == 0:
            xp = x
        elif xp < x:
            ys, ye = 0, h - 1
        xs, xe, ys, ye = min(xs, x), max(xe, x), min(ys, y), max(ye, y)
    return [float(xs), float(ys), float(xe - xs + 1), float(ye - ys + 1)]


def rle_decode(counts, h: int, w: int) -> numpy.ndarray:
    """`rleDecode`: uint8 [h, w] mask (column-major runs)."""
    flat = numpy.zeros(h * w, dtype=numpy.uint8)
    pos, val = 0, 0
    for c in counts:
        c = int(c)
        if val:
            flat[pos:pos + c] = 1
        pos += c
        val ^= 1
    return flat.reshape(w, h).T.copy()


def fr_py_objects(polygons: List[numpy.ndarray], h: int, w: int) -> List[dict]:
    """`pycocotools.mask.frPyObjects` for a list of polygons: [{'size': [h, w], 'counts': bytes}, ...]."""
    return [{'size': [int(h), int(w)], 'counts': rle_to_string(rle_from_polygon(p, h, w))} for p in polygons]


# ------------------------------------------------------------------------------------------- COCOGtCreator
class COCOGtCreator:
    """coco_gt.py:14-134, method for method."""

    def __init__(self, class_to_color_map: Dict, image_root: Path = Path('/')):
        self.class_to_color_map = class_to_color_map
        self.categories = self.build_categories()
        self.image_root = image_root

    def build_categories(self) -> List[dict]:
        return [{'id': category_id, 'name': class_name, 'supercategory': class_name, 'color': color}
                for category_id, (class_name, color) in enumerate(self.class_to_color_map.items())]

    @staticmethod
    def get_label_image(image_data) -> numpy.ndarray:
        image_data = numpy.array(image_data)
        _, label_image = numpy.split(image_data, 2, axis=1)
        return label_image

    @staticmethod
    def extract_rle(class_mask: numpy.ndarray) -> List[dict]:
        """:35-49: external contours (CHAIN_APPROX_SIMPLE) with at least 3 points, each encoded as a polygon RLE."""
        contours, _ = cv2.findContours(class_mask.astype('uint8'), cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
        polygons = [contour.ravel() for contour in contours if contour.size >= 6]
        if len(polygons) == 0:
            return polygons
        return fr_py_objects(polygons, class_mask.shape[-2], class_mask.shape[-1])

    @staticmethod
    def _rgb(color):
        return tuple(color) if not isinstance(color, str) else ImageColor.getrgb(color)

    def determine_classes_in_image(self, image_data) -> Dict[str, bool]:
        label_image = self.get_label_image(image_data)
        classes_in_image = {}
        for class_name, color in self.class_to_color_map.items():
            if class_name == 'background':
                continue
            class_mask = numpy.multiply.reduce(label_image[:, :] == self._rgb(color), axis=2)
            classes_in_image[f'has_{class_name}'] = len(self.extract_rle(class_mask)) > 0
        return classes_in_image

    def build_annotations_for_image(self, image_data, image_id: int, annotation_id: int) -> Tuple[List[dict], int]:
        label_image = self.get_label_image(image_data)
        annotations = []
        for class_id, (class_name, color) in enumerate(self.class_to_color_map.items()):
            if class_name == 'background':
                continue                                   # no need to annotate background
            class_mask = numpy.multiply.reduce(label_image[:, :] == self._rgb(color), axis=2)
            for rle in self.extract_rle(class_mask):
                h, w = rle['size']
                counts = rle_from_string(rle['counts'])
                rle['counts'] = rle['counts'].decode('utf-8')
                annotations.append({'id': annotation_id, 'image_id': image_id, 'category_id': class_id, 'segmentation': rle,
                                    'area': rle_area(counts), 'bbox': rle_to_bbox(counts, h, w), 'iscrowd': 0})
                annotation_id += 1
        return annotations, annotation_id

    def create_coco_gt_from_image_paths(self, image_paths: Iterable[Path]) -> dict:
        images, annotations = [], []
        annotation_id = 0
        for i, image_path in enumerate(image_paths):
            with Image.open(str(image_path)) as the_image:
                images.append({'id': i, 'width': the_image.width // 2, 'height': the_image.height,
                               'file_name': str(Path(image_path).relative_to(self.image_root)), 'license': 0, 'flickr_url': '',
                               'coco_url': '', 'date_captured': str(datetime.datetime.utcnow())})
                annotations_for_image, annotation_id = self.build_annotations_for_image(the_image, i, annotation_id)
                annotations.extend(annotations_for_image)
        return {
            'info': {'year': datetime.date.today().year, 'version': '1',
                     'description': 'COCO GT for evaluation of semantic segmentation', 'contributor': 'yourself',
                     'url': 'http://example.com'},
            'images': images, 'annotations': annotations, 'categories': self.categories,
            'licenses': [{'id': 0, 'name': 'Kekse', 'url': 'http://example.com'}],
        }


def iter_through_images_in(image_root: Path, extension: str = 'png') -> Iterable[Path]:
    """coco_gt.py:137-140 (glob order, as the reference)."""
    yield from Path(image_root).glob(f'**/*.{extension}')
