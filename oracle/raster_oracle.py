"""ORACLE / TEST INFRASTRUCTURE ONLY — CPU restatement of the IMAGE branch of DynEnv's getAgentVision.

Draws the 8-channel (top camera 0..3, bottom camera 4..7) 480 x 640 observation of one robot from the raw
detection lists of the C oracle (oracle/rc_oracle.c, RcRawVision), following the reference line by line:
  RoboCupEnvironment.py:1318-1467,1533-1534   the drawing calls and their order
  cutils.py:166-200                            pin-hole camera matrices
  cutils.py:206-222 projectPoints, :226-257 getConicPoints, :261-284 estimateConic
The rasterisation itself is OpenCV's (cv2.line / polylines / rectangle / circle on float64 planes), the very
library the reference calls; it is present in this image and on the GPU box.  Pinned against the reference's own
reset() in IMAGE mode by tests/golden/robocup_image_*.npz (tests/test_oracle_golden_image.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this module.
"""
import math

import cv2
import numpy as np

LINE_WIDTH = 5.0  # RoboCupEnvironment.lineWidth (:80)

# ---- cutils.py:166-200
focal = 543.6
A = np.array([[focal, 0, 319.5], [0, -focal, 239.5], [0, 0, 1]])


def _cam(ang, ty, tz):
    rot = np.array([[1, 0, 0], [0, np.cos(ang), -np.sin(ang)], [0, np.sin(ang), np.cos(ang)]])
    tr = np.concatenate((np.concatenate((rot, np.array([[0, ], [ty, ], [tz, ]])), axis=1), np.array([[0, 0, 0, 1, ], ])), axis=0)
    return np.matmul(A, np.linalg.inv(tr)[:3])


bottomTr = _cam(0.6929 + 0.25, 53.774, 5.071)
topTr = _cam(0.0209 + 0.25, 58.364, 5.871)


def projectPoints(points, compRadius=True):
    """cutils.py:206-222"""
    topProj = np.matmul(topTr, points)
    bottomProj = np.matmul(bottomTr, points)
    topProj = topProj[0:2] / topProj[2]
    bottomProj = bottomProj[0:2] / bottomProj[2]
    tRad = 0
    bRad = 0
    if compRadius:
        tRad = np.sqrt(np.sum(np.square(topProj[:, 0] - topProj[:, 1])))
        bRad = np.sqrt(np.sum(np.square(bottomProj[:, 0] - bottomProj[:, 1])))
    return topProj, math.ceil(tRad), bottomProj, math.ceil(bRad)


def getConicPoints(yRange, center, params):
    """cutils.py:226-257"""
    yCoord = np.arange(0, yRange) - center[1]
    a = params[0]
    a4 = 4 * a
    overa = 1.0 / (2 * a)
    b = yCoord * params[2] + params[3]
    c = yCoord * (yCoord * params[1] + params[4]) + params[5]
    sqr = b * b - a4 * c
    ind = sqr >= 0
    sqrt = np.sqrt(sqr[ind])
    x1 = np.round((-b[ind] + sqrt) * overa + center[0])
    x2 = np.round((-b[ind] - sqrt) * overa + center[0])
    yCoord = np.round(yCoord + center[1])
    c1 = np.vstack((x1, yCoord[ind])).astype('int32').transpose()
    c2 = np.vstack((x2, yCoord[ind])).astype('int32').transpose()
    return c1, c2


def estimateConic(points):
    """cutils.py:261-284: null vector of the 8 x 6 design matrix [x^2, y^2, xy, x, y, 1]"""
    points = points.transpose()
    X = np.array([[p[0] * p[0], p[1] * p[1], p[0] * p[1], p[0], p[1], 1] for p in points[:8]])
    u, s, vt = np.linalg.svd(X)
    return -vt[-1]


def render(raw, line_width=LINE_WIDTH):
    """raw: dict of one RcRawVision (oracle.RcOracle.raw_vision).  Returns float64 [8, 480, 640] like the reference
    (RoboCupEnvironment.py:1318-1467, :1533-1534)."""
    bottomCamImg = np.zeros((4, 480, 640))
    topCamImg = np.zeros((4, 480, 640))
    cnt = raw["counts"]
    for q in range(int(cnt[5])):  # lines (:1324-1341)
        x1, y1, x2, y2 = [float(v) for v in raw["line"][q]]
        linevec = np.array([
            [(-y1 - y2) / 2, 0, (x1 + x2) / 2, 1],
            [(-y1 - y2) / 2 + line_width / 2, 0, (x1 + x2) / 2, 1],
            [-y1, 0, x1, 1],
            [-y2, 0, x2, 1]]).transpose()
        tProj, tRad, bProj, bRad = projectPoints(linevec)
        cv2.line(topCamImg[3], (int(tProj[0, 2]), int(tProj[1, 2])), (int(tProj[0, 3]), int(tProj[1, 3])), 1, tRad)
        cv2.line(bottomCamImg[3], (int(bProj[0, 2]), int(bProj[1, 2])), (int(bProj[0, 3]), int(bProj[1, 3])), 1, bRad)
    if raw["circle"][0] != 0:  # centre circle (:1343-1393)
        cx, cy, cr = [float(v) for v in raw["circle"][1:4]]
        eo = (cr * math.cos(math.pi / 4), cr * math.sin(math.pi / 4))  # Vec2d(r, 0).rotated(pi / 4)
        circlevec = np.array([
            [-cy, 0, cx, 1],
            [-cy, 0, cx - cr, 1],
            [-cy, 0, cx + cr, 1],
            [-cy - cr, 0, cx, 1],
            [-cy + cr, 0, cx, 1],
            [-cy - eo[1], 0, cx + eo[0], 1],
            [-cy + eo[0], 0, cx - eo[1], 1],
            [-cy + eo[1], 0, cx + eo[0], 1],
            [-cy - eo[0], 0, cx - eo[1], 1]]).transpose()
        tProj, tRad, bProj, bRad = projectPoints(circlevec, False)
        length = math.sqrt(cx * cx + cy * cy)
        tThickness = 15 - max(0, min(14, int(length / 40)))
        bThickness = 30 - max(0, min(29, int(length / 20)))
        tParams = estimateConic(tProj[:, 1:] - tProj[:, 0:1])
        bParams = estimateConic(bProj[:, 1:] - bProj[:, 0:1])
        tx1, tx2 = getConicPoints(480, tProj[:, 0], tParams)
        bx1, bx2 = getConicPoints(480, bProj[:, 0], bParams)
        cv2.polylines(topCamImg[3], [tx1], False, 1, tThickness)
        cv2.polylines(topCamImg[3], [tx2], False, 1, tThickness)
        cv2.polylines(bottomCamImg[3], [bx1], False, 1, bThickness)
        cv2.polylines(bottomCamImg[3], [bx2], False, 1, bThickness)
        if tx1.shape[0]:
            if tx1[0, 1]:
                cv2.line(topCamImg[3], tuple(int(v) for v in tx1[0]), tuple(int(v) for v in tx2[0]), 1, tThickness)
            if tx1[-1, 1] < 480 - 1:
                cv2.line(topCamImg[3], tuple(int(v) for v in tx1[-1]), tuple(int(v) for v in tx2[-1]), 1, tThickness)
        if bx1.shape[0]:
            if bx1[0, 1]:
                cv2.line(bottomCamImg[3], tuple(int(v) for v in bx1[0]), tuple(int(v) for v in bx2[0]), 1, bThickness)
            if bx1[-1, 1] < 480 - 1:
                cv2.line(bottomCamImg[3], tuple(int(v) for v in bx1[-1]), tuple(int(v) for v in bx2[-1]), 1, bThickness)
    for k in range(int(cnt[1])):  # robots: filled rectangles (:1395-1408)
        x, y, r = [float(v) for v in raw["robot"][k]]
        robvec = np.array([[-y - r, 0, x, 1], [-y + r, 58, x, 1]]).transpose()
        tProj, tRad, bProj, bRad = projectPoints(robvec, False)
        cv2.rectangle(topCamImg[1], (int(tProj[0, 0]), int(tProj[1, 0])), (int(tProj[0, 1]), int(tProj[1, 1])), 1, -1)
        cv2.rectangle(bottomCamImg[1], (int(bProj[0, 0]), int(bProj[1, 0])), (int(bProj[0, 1]), int(bProj[1, 1])), 1, -1)
    for k in range(int(cnt[2])):  # goal posts: vertical thick lines (:1410-1425)
        x, y, r = [float(v) for v in raw["post"][k]]
        goalvec = np.array([[-y, 0, x, 1], [-y + r / 2, 0, x, 1], [-y, 80, x, 1]]).transpose()
        tProj, tRad, bProj, bRad = projectPoints(goalvec)
        cv2.line(topCamImg[2], (int(tProj[0, 0]), int(tProj[1, 0])), (int(tProj[0, 2]), int(tProj[1, 2])), 1, tRad)
        cv2.line(bottomCamImg[2], (int(bProj[0, 0]), int(bProj[1, 0])), (int(bProj[0, 2]), int(bProj[1, 2])), 1, bRad)
    for name, n in (("cross", int(cnt[3])), ("fcross", int(cnt[4]))):  # crosses: filled circles (:1427-1451)
        for k in range(n):
            x, y, r = [float(v) for v in raw[name][k]]
            crossvec = np.array([[-y, 0, x, 1], [-y + r / 2, 0, x, 1]]).transpose()
            tProj, tRad, bProj, bRad = projectPoints(crossvec)
            cv2.circle(topCamImg[3], (int(tProj[0, 0]), int(tProj[1, 0])), tRad, 1, -1)
            cv2.circle(bottomCamImg[3], (int(bProj[0, 0]), int(bProj[1, 0])), bRad, 1, -1)
    for k in range(int(cnt[0])):  # balls (:1453-1466)
        x, y, r = [float(v) for v in raw["ball"][k]]
        ballvec = np.array([[-y, r / 2, x, 1], [-y + r / 2, r / 2, x, 1]]).transpose()
        tProj, tRad, bProj, bRad = projectPoints(ballvec)
        cv2.circle(topCamImg[0], (int(tProj[0, 0]), int(tProj[1, 0])), tRad, 1, -1)
        cv2.circle(bottomCamImg[0], (int(bProj[0, 0]), int(bProj[1, 0])), bRad, 1, -1)
    return np.concatenate((topCamImg, bottomCamImg))


def pack(img):
    """[8, 480, 640] 0/1 image -> uint32 [8, 480, 20], bit x % 32 of word x // 32 = pixel x (the C-ABI layout)"""
    bits = (np.asarray(img) != 0).astype(np.uint8).reshape(8, 480, 20, 32)
    return (bits.astype(np.uint64) << np.arange(32, dtype=np.uint64)).sum(-1).astype(np.uint32)


def unpack(words):
    w = np.asarray(words, dtype=np.uint32)
    return ((w[..., None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(w.shape[:-1] + (w.shape[-1] * 32,)).astype(np.uint8)
