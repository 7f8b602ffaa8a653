"""CPU: the restated COCO mask API (polygon -> RLE, string codec, area, bbox) and COCOGtCreator.

pycocotools 2.0.2 (the reference's dependency) is not installed here, so there are no golden vectors: the restatement
is pinned through identities of the published API -- `rleFrBbox` is `rleFrPoly` on the box's four corners and must give
area w*h and decode to exactly that box; the string codec must round-trip (incl. negative deltas); decode(encode(polygon))
must lie within one pixel of OpenCV's rasterisation of the polygon."""
import json

import cv2
import numpy as np
import pytest
from PIL import Image

from synthesis_in_style_b200 import coco_gt as cg

COLORS = {'background': '#000000', 'printed_text': '#0000FF', 'handwritten_text': '#FF0000'}


@pytest.mark.parametrize('box', [(3, 4, 10, 7), (0, 0, 50, 40), (10, 5, 1, 1), (49, 39, 1, 1), (0, 12, 50, 3)])
def test_box_polygon_identities(box):
    h, w = 40, 50
    x0, y0, bw, bh = box
    counts = cg.rle_from_polygon([x0, y0, x0, y0 + bh, x0 + bw, y0 + bh, x0 + bw, y0], h, w)     # rleFrBbox's corner order
    mask = cg.rle_decode(counts, h, w)
    want = np.zeros((h, w), np.uint8)
    want[y0:y0 + bh, x0:x0 + bw] = 1
    assert np.array_equal(mask, want)
    assert cg.rle_area(counts) == bw * bh
    assert cg.rle_to_bbox(counts, h, w) == [float(x0), float(y0), float(bw), float(bh)]
    assert int(counts.sum()) == h * w


def test_string_codec_round_trip():
    rng = np.random.RandomState(0)
    for _ in range(50):
        counts = rng.randint(0, 5000, size=rng.randint(1, 40)).tolist()
        s = cg.rle_to_string(counts)
        assert all(48 <= b <= 111 for b in s)
        assert cg.rle_from_string(s) == counts
        assert cg.rle_from_string(s.decode('ascii')) == counts
    assert cg.rle_to_string([0]) == b'0' and cg.rle_to_string([5, 3, 5]) == b'535'
    assert cg.rle_from_string(cg.rle_to_string([100, 2, 100, 1])) == [100, 2, 100, 1]    # 4th count stored as -1


def test_polygon_between_open_and_closed_raster():
    h, w = 64, 64
    rng = np.random.RandomState(1)
    for _ in range(20):
        canvas = np.zeros((h, w), np.uint8)
        cv2.ellipse(canvas, (int(rng.randint(15, 50)), int(rng.randint(15, 50))), (int(rng.randint(4, 14)), int(rng.randint(4, 14))),
                    float(rng.randint(0, 180)), 0, 360, 1, -1)
        contours, _ = cv2.findContours(canvas, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
        poly = contours[0].ravel()
        mask = cg.rle_decode(cg.rle_from_polygon(poly, h, w), h, w)
        closed = np.zeros((h, w), np.uint8)
        cv2.fillPoly(closed, [contours[0]], 1)                       # includes the boundary pixels
        interior = cv2.erode(closed, np.ones((3, 3), np.uint8))
        outer = cv2.dilate(closed, np.ones((3, 3), np.uint8))
        # the API's raster treats pixel i as [i, i+1) and rounds the traced boundary: it stays within one pixel of OpenCV's
        assert (mask <= outer).all() and (interior <= mask).all()
        assert 0.6 * closed.sum() <= mask.sum() <= closed.sum()


def test_coco_gt_creator_structure(tmp_path):
    label = np.zeros((32, 32, 3), np.uint8)
    label[4:12, 5:20] = (0, 0, 255)
    label[20:28, 8:16] = (255, 0, 0)
    label[30, 30] = (255, 0, 0)                       # a single pixel: contour of one point, skipped (size < 6)
    image = np.concatenate([np.full((32, 32, 3), 127, np.uint8), label], axis=1)
    (tmp_path / '0' / '0').mkdir(parents=True)
    path = tmp_path / '0' / '0' / '0000.png'
    Image.fromarray(image).save(path)
    creator = cg.COCOGtCreator(COLORS, image_root=tmp_path)
    assert creator.categories[1] == {'id': 1, 'name': 'printed_text', 'supercategory': 'printed_text', 'color': '#0000FF'}
    with Image.open(path) as im:
        assert creator.determine_classes_in_image(im) == {'has_printed_text': True, 'has_handwritten_text': True}
    gt = creator.create_coco_gt_from_image_paths([path])
    json.dumps(gt)                                   # serialisable as the reference dumps it
    assert [im['file_name'] for im in gt['images']] == ['0/0/0000.png'] and gt['images'][0]['width'] == 32
    anns = gt['annotations']
    assert [a['category_id'] for a in anns] == [1, 2] and [a['id'] for a in anns] == [0, 1]
    a0 = anns[0]
    assert a0['segmentation']['size'] == [32, 32] and isinstance(a0['segmentation']['counts'], str)
    # contour points are pixel centres: the 15 x 8 block yields the polygon (5,4)-(19,11) -> 14 x 7 pixels by the centre rule
    assert a0['bbox'] == [5.0, 4.0, 14.0, 7.0] and a0['area'] == 14 * 7 and a0['iscrowd'] == 0
    mask = cg.rle_decode(cg.rle_from_string(a0['segmentation']['counts']), 32, 32)
    assert mask[4:11, 5:19].all() and mask.sum() == a0['area']
