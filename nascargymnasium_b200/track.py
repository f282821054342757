"""Host-side track ingestion: ``.track`` text -> segments -> wall boxes -> device tables.

Restates (independently, float64, same operation order so the numbers agree
bit for bit with the reference's):

* the ``.track`` parser            /root/reference/src/track_generator.py:305-405
* segment geometry                 /root/reference/src/track_generator.py:66-174
* wall-box construction            /root/reference/src/car_physics.py:118-339
* the float64 -> float32 hand-over Box2D performs when a static body is
  created (position, angle, ``sinf``/``cosf`` of the float32 angle, half extents).

On top of that it builds what only the B200 engine needs: a uniform grid over
the wall boxes for the ray cull, the fat AABBs for the contact broad phase and
one contiguous, 16-byte aligned *track blob* that a CTA stages into shared
memory with a single ``cp.async.bulk`` (see DESIGN.md "Data layout").
"""
from __future__ import annotations

import ctypes
import ctypes.util
import math
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import constants as K

# ----------------------------------------------------------------------------
# Built-in tracks.  These are the geometric parameters of the eight ovals the
# reference ships as tracks/*.track (width, then STRAIGHT length [bank] /
# LEFT|RIGHT angle radius [bank]); every track starts GRID, STARTLINE.
# ----------------------------------------------------------------------------
BUILTIN_TRACKS: Dict[str, Tuple[float, Tuple[Tuple, ...]]] = {
    "daytona": (12, (("S", 1000), ("L", 180, 260, 31), ("S", 1200), ("L", 180, 260, 31), ("S", 95))),
    "martinsville": (16, (("S", 15), ("L", 180, 89, 12), ("S", 145), ("L", 180, 89, 12), ("S", 25))),
    "michigan": (22, (("S", 383, 5), ("L", 160, 290, 18), ("L", 40, 1297, 12), ("L", 160, 290, 18), ("S", 200, 5))),
    "nascar": (25, (("S", 200), ("L", 180, 300), ("S", 400), ("L", 180, 300), ("S", 95))),
    "nascar2": (25, (("S", 200), ("L", 90, 160, 31), ("L", 90, 250, 28), ("S", 400), ("L", 80, 140, 30),
                     ("L", 100, 251, 23), ("S", 76))),
    "nascar_banked": (25, (("S", 200, 0), ("L", 180, 300, 31), ("S", 400, 0), ("L", 180, 300, 31), ("S", 95, 0))),
    "talladega": (15, (("S", 1014), ("L", 160, 285, 33), ("S", 475), ("L", 40, 761, 17), ("S", 475),
                       ("L", 160, 285, 33), ("S", 100))),
    "trioval": (12, (("S", 970), ("L", 160, 260), ("S", 541), ("L", 40, 150), ("S", 774), ("L", 160, 219),
                     ("S", 100))),
}
# sorted file-name order, i.e. what glob()+sort gives over tracks/*.track
BUILTIN_TRACK_NAMES = tuple(sorted(BUILTIN_TRACKS))


def builtin_track_text(name: str) -> str:
    """The ``.track`` text of a built-in track (same grammar the parser reads)."""
    width, cmds = BUILTIN_TRACKS[name]
    out = [f"WIDTH {width:g}", "GRID", "STARTLINE"]
    for c in cmds:
        if c[0] == "S":
            out.append("STRAIGHT " + " ".join(f"{v:g}" for v in c[1:]))
        else:
            out.append(("LEFT " if c[0] == "L" else "RIGHT ") + " ".join(f"{v:g}" for v in c[1:]))
    return "\n".join(out) + "\n"


@dataclass
class Segment:
    """One track segment (track_generator.py:12-41)."""
    kind: str                 # GRID | STARTLINE | STRAIGHT | FINISHLINE | CURVE
    length: float             # metres (arc length for curves)
    start: Tuple[float, float]
    end: Tuple[float, float]
    width: float
    curve_angle: float = 0.0
    curve_radius: float = 0.0
    curve_direction: str = ""
    start_heading: float = 0.0
    end_heading: float = 0.0
    banking: float = 0.0


@dataclass
class Track:
    """Parsed track: ordered segments + arc length (track_generator.py:43-174)."""
    name: str = ""
    width: float = K.DEFAULT_TRACK_WIDTH
    segments: List[Segment] = field(default_factory=list)
    total_length: float = 0.0
    _pos: Tuple[float, float] = (0.0, 0.0)
    _heading: float = 0.0

    def add_segment(self, kind: str, length: float, curve_angle: float = 0.0, curve_radius: float = 0.0,
                    curve_direction: str = "", banking: float = 0.0) -> None:
        start, h0 = self._pos, self._heading
        if curve_angle == 0.0:
            h1 = h0
            hr = math.radians(h0)
            end = (start[0] + length * math.cos(hr), start[1] + length * math.sin(hr))
        else:
            kind = "CURVE"
            h0r = math.radians(h0)
            car = math.radians(curve_angle)
            tm = 1.0 if curve_direction == "LEFT" else -1.0
            perp = h0r + tm * math.pi / 2
            cx = start[0] + curve_radius * math.cos(perp)
            cy = start[1] + curve_radius * math.sin(perp)
            h1 = h0 + tm * curve_angle
            a0 = h0r - tm * math.pi / 2
            a1 = a0 + tm * car
            end = (cx + curve_radius * math.cos(a1), cy + curve_radius * math.sin(a1))
            length = abs(curve_radius * math.radians(curve_angle))
        self.segments.append(Segment(kind, length, start, end, self.width, curve_angle, curve_radius,
                                     curve_direction, h0, h1, banking))
        self.total_length += length
        self._pos, self._heading = end, h1

    @property
    def has_banking(self) -> bool:
        """car_physics.py:674-690."""
        return any(abs(s.banking) >= K.BANKING_MIN_ANGLE for s in self.segments)

    @property
    def startline(self) -> Optional[Segment]:
        for s in self.segments:
            if s.kind == "STARTLINE":
                return s
        return None


def parse_track_text(text: str, name: str = "") -> Track:
    """The reference's line grammar (track_generator.py:326-403), same errors."""
    track = Track(name=name)
    for raw in text.splitlines():
        line = raw.strip().upper()
        if not line or line.startswith("#"):
            continue
        if "#" in line:
            line = line[: line.find("#")].strip()
        parts = line.split()
        if not parts:
            continue
        cmd = parts[0]
        if cmd == "WIDTH":
            if len(parts) != 2:
                raise ValueError(f"WIDTH command requires exactly one argument: {line}")
            try:
                track.width = float(parts[1])
            except ValueError:
                raise ValueError(f"Invalid width value: {parts[1]}")
        elif cmd == "GRID":
            track.add_segment("GRID", K.GRID_LENGTH)
        elif cmd == "STARTLINE":
            track.add_segment("STARTLINE", K.STARTLINE_LENGTH)
        elif cmd == "FINISHLINE":
            track.add_segment("FINISHLINE", K.FINISHLINE_LENGTH)
        elif cmd == "STRAIGHT":
            if len(parts) < 2 or len(parts) > 3:
                raise ValueError(f"STRAIGHT command requires 1-2 arguments (length [, banking]): {line}")
            try:
                length = float(parts[1])
                banking = float(parts[2]) if len(parts) == 3 else 0.0
            except ValueError:
                raise ValueError(f"Invalid numeric values for STRAIGHT command: {parts[1:]}")
            if banking < -45 or banking > 45:
                raise ValueError(f"Banking angle must be between -45 and 45 degrees: {banking}")
            track.add_segment("STRAIGHT", length, banking=banking)
        elif cmd in ("LEFT", "RIGHT"):
            if len(parts) < 3 or len(parts) > 4:
                raise ValueError(f"{cmd} command requires 2-3 arguments (angle, radius [, banking]): {line}")
            try:
                angle = float(parts[1])
                radius = float(parts[2])
                banking = float(parts[3]) if len(parts) == 4 else 0.0
            except ValueError:
                raise ValueError(f"Invalid numeric values for {cmd} command: {parts[1:]}")
            if angle <= 0 or angle > 360:
                raise ValueError(f"Curve angle must be between 0 and 360 degrees: {angle}")
            if radius <= 0:
                raise ValueError(f"Curve radius must be positive: {radius}")
            if banking < -45 or banking > 45:
                raise ValueError(f"Banking angle must be between -45 and 45 degrees: {banking}")
            track.add_segment("CURVE", 0, curve_angle=angle, curve_radius=radius, curve_direction=cmd,
                              banking=banking)
        else:
            raise ValueError(f"Unknown command: {cmd}")
    return track


def track_name_of(path: str) -> str:
    return os.path.splitext(os.path.basename(path))[0]


def load_track(path_or_name: str) -> Track:
    """Load a ``.track`` file; a missing file whose stem names a built-in oval
    (``tracks/daytona.track``) resolves to the built-in definition, otherwise
    ``FileNotFoundError`` exactly like ``TrackLoader.load_track`` (:318)."""
    if os.path.exists(path_or_name):
        with open(path_or_name, "r") as f:
            return parse_track_text(f.read(), track_name_of(path_or_name))
    stem = track_name_of(path_or_name)
    if stem in BUILTIN_TRACKS:
        return parse_track_text(builtin_track_text(stem), stem)
    raise FileNotFoundError(f"Track file not found: {path_or_name}")


# ----------------------------------------------------------------------------
# Wall boxes (car_physics.py:118-339), float64 stage
# ----------------------------------------------------------------------------
def wall_lines(track: Track) -> np.ndarray:
    """(n,4) float64 [x1,y1,x2,y2] wall centre lines in creation order."""
    lines: List[Tuple[float, float, float, float]] = []

    def add(x1, y1, x2, y2):
        length = ((x2 - x1) ** 2 + (y2 - y1) ** 2) ** 0.5
        if length < 0.1:
            return
        lines.append((x1, y1, x2, y2))

    for seg in track.segments:
        hw = seg.width / 2
        if seg.kind == "CURVE":
            if seg.curve_radius <= 0 or seg.curve_angle <= 0:
                continue
            h0 = math.radians(seg.start_heading)
            ca = math.radians(seg.curve_angle)
            tm = 1.0 if seg.curve_direction == "LEFT" else -1.0
            perp = h0 + tm * math.pi / 2
            cx = seg.start[0] + seg.curve_radius * math.cos(perp)
            cy = seg.start[1] + seg.curve_radius * math.sin(perp)
            if seg.curve_direction == "LEFT":
                r_in, r_out = seg.curve_radius - hw, seg.curve_radius + hw
            else:
                r_in, r_out = seg.curve_radius + hw, seg.curve_radius - hw
            a0 = h0 - tm * math.pi / 2
            n = max(K.CURVE_MIN_SEGMENTS,
                    min(K.CURVE_MAX_SEGMENTS, int(abs(seg.curve_angle) / K.CURVE_DEGREES_PER_SEGMENT)))
            inner, outer = [], []
            for i in range(n + 1):
                t = i / n
                a = a0 + tm * ca * t
                if r_in > 0:
                    inner.append((cx + r_in * math.cos(a), cy + r_in * math.sin(a)))
                outer.append((cx + r_out * math.cos(a), cy + r_out * math.sin(a)))
            if len(inner) < 2 or len(outer) < 2:
                continue
            for i in range(len(inner) - 1):
                add(inner[i][0], inner[i][1], inner[i + 1][0], inner[i + 1][1])
            for i in range(len(outer) - 1):
                add(outer[i][0], outer[i][1], outer[i + 1][0], outer[i + 1][1])
        else:
            sx, sy = seg.start
            ex, ey = seg.end
            sl = math.sqrt((ex - sx) ** 2 + (ey - sy) ** 2)
            if sl > 0:
                dx, dy = (ex - sx) / sl, (ey - sy) / sl
                px, py = -dy, dx
            else:
                px, py = 0, 1
            add(sx + px * hw, sy + py * hw, ex + px * hw, ey + py * hw)
            add(sx - px * hw, sy - py * hw, ex - px * hw, ey - py * hw)
    return np.asarray(lines, dtype=np.float64).reshape(-1, 4)


_libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
_libm.sinf.restype = ctypes.c_float
_libm.sinf.argtypes = [ctypes.c_float]
_libm.cosf.restype = ctypes.c_float
_libm.cosf.argtypes = [ctypes.c_float]


def wall_boxes(track: Track) -> np.ndarray:
    """(n,7) float32 [px,py,c,s,hx,hy,angle]: what Box2D stores for each static wall
    body (``_create_wall_body_from_line`` :293-315 then b2Body/b2Rot float32)."""
    lines = wall_lines(track)
    out = np.zeros((len(lines), 7), dtype=np.float32)
    for i, (x1, y1, x2, y2) in enumerate(lines):
        cx = (x1 + x2) / 2
        cy = (y1 + y2) / 2
        length = ((x2 - x1) ** 2 + (y2 - y1) ** 2) ** 0.5
        angle = np.float32(math.atan2(y2 - y1, x2 - x1))
        out[i, 0] = np.float32(cx)
        out[i, 1] = np.float32(cy)
        out[i, 2] = _libm.cosf(float(angle))
        out[i, 3] = _libm.sinf(float(angle))
        out[i, 4] = np.float32(length / 2)
        out[i, 5] = np.float32(K.WALL_THICKNESS / 2)
        out[i, 6] = angle
    return out


# ----------------------------------------------------------------------------
# Device tables
# ----------------------------------------------------------------------------
MAX_SEGS = 16
SEG_STRIDE = 12          # floats per segment row
WALL_STRIDE = 8          # floats per wall row  [px,py,c,s,hx for rays,hy,angle,hx]
HDR_WORDS = 32
GRID_CELL = 32.0
POLY_RADIUS = 0.01       # b2_polygonRadius
AABB_EXT = 0.1           # b2_aabbExtension

# header word indices (ints are stored bit-cast in the float32 blob)
H_NWALLS, H_NSEGS, H_GNX, H_GNY, H_HASBANK, H_WORDS, H_OFF_SEGS, H_OFF_WALLS, H_OFF_AABB, H_OFF_CELLS, \
    H_OFF_ITEMS, H_NITEMS = range(12)
H_GX0, H_GY0, H_INVCELL, H_CELL, H_LTOT, H_MINLAP, H_SLX0, H_SLY0, H_SLDX, H_SLDY, H_SLLEN2, H_SLHALFW, \
    H_HALF_LTOT = range(12, 25)
H_STAGE_WORDS = 25       # words [0, H_STAGE_WORDS) are what a CTA stages into shared memory (the whole table)
H_OFF_SEG64 = 26         # float64 rows [sx,sy,ex,ey,cum_chord] per segment (tie-exact nearest-segment search)
H_OFF_SEGMASK = 27       # uint16 per grid cell: which segments can be the chord-nearest one for a point of that cell
ITEM_BLOCK = 4           # grid item lists are padded to blocks of 4 (one 8-byte load, four slab tests in flight)
SEG64_STRIDE = 5


def _wall_corners(b: np.ndarray) -> np.ndarray:
    """(n,4,2) float32 world corners, computed like b2Mul(xf, v) in float32."""
    px, py, c, s, hx, hy = [b[:, i].astype(np.float32) for i in range(6)]
    vx = np.stack([-hx, hx, hx, -hx], axis=1)
    vy = np.stack([-hy, -hy, hy, hy], axis=1)
    x = (c[:, None] * vx - s[:, None] * vy) + px[:, None]
    y = (s[:, None] * vx + c[:, None] * vy) + py[:, None]
    return np.stack([x, y], axis=2).astype(np.float32)


def wall_fat_aabbs(b: np.ndarray) -> np.ndarray:
    """(n,4) float32 [lx,ly,ux,uy]: b2PolygonShape::ComputeAABB (+-radius) then
    the tree's +-b2_aabbExtension, as stored at proxy creation."""
    cs = _wall_corners(b)
    r = np.float32(POLY_RADIUS)
    e = np.float32(AABB_EXT)
    lo = cs.min(axis=1) - r - e
    hi = cs.max(axis=1) + r + e
    return np.concatenate([lo, hi], axis=1).astype(np.float32)


def _obb_overlaps_cell(bx, cell_lo, cell_hi, margin):
    """SAT test: oriented wall box (inflated by margin) vs axis-aligned cell."""
    px, py, c, s, hx, hy = [float(v) for v in bx[:6]]
    hx += margin
    hy += margin
    ccx, ccy = 0.5 * (cell_lo[0] + cell_hi[0]), 0.5 * (cell_lo[1] + cell_hi[1])
    chx, chy = 0.5 * (cell_hi[0] - cell_lo[0]), 0.5 * (cell_hi[1] - cell_lo[1])
    dx, dy = px - ccx, py - ccy
    # world axes
    if abs(dx) > chx + abs(c) * hx + abs(s) * hy:
        return False
    if abs(dy) > chy + abs(s) * hx + abs(c) * hy:
        return False
    # box axes
    if abs(dx * c + dy * s) > hx + chx * abs(c) + chy * abs(s):
        return False
    if abs(-dx * s + dy * c) > hy + chx * abs(s) + chy * abs(c):
        return False
    return True


@dataclass
class TrackTable:
    """Everything the engine (and the tests) need about one track."""
    track: Track
    boxes: np.ndarray        # (n,7) f32 [px,py,c,s,hx,hy,angle]
    fat_aabb: np.ndarray     # (n,4) f32
    segs: np.ndarray         # (MAX_SEGS, SEG_STRIDE) f32
    seg64: np.ndarray        # (nseg, 6) f64 [sx,sy,ex,ey,banking,0] for the oracle
    grid_origin: Tuple[float, float]
    grid_dims: Tuple[int, int]
    cell_start: np.ndarray   # (ncells+1,) uint16
    cell_items: np.ndarray   # (nitems,) uint16
    blob: np.ndarray         # (words,) float32, 16-byte multiple

    @property
    def n_walls(self) -> int:
        return int(self.boxes.shape[0])


def build_track_table(track: Track, cell: float = GRID_CELL) -> TrackTable:
    boxes = wall_boxes(track)
    n = len(boxes)
    if n == 0:
        raise ValueError("track has no walls")
    if n >= 32768:
        raise ValueError("too many wall boxes for 15-bit contact indices")
    segs_src = track.segments
    if len(segs_src) > MAX_SEGS:
        raise ValueError(f"track has {len(segs_src)} segments; this engine supports at most {MAX_SEGS}")
    fat = wall_fat_aabbs(boxes)

    # --- segment rows (progress / banking / start line use the chords) -----
    segs = np.zeros((MAX_SEGS, SEG_STRIDE), dtype=np.float32)
    seg64 = np.zeros((len(segs_src), 6), dtype=np.float64)
    cum = 0.0
    for i, s in enumerate(segs_src):
        sx, sy = s.start
        ex, ey = s.end
        dx, dy = ex - sx, ey - sy
        l2 = dx * dx + dy * dy
        chord = math.sqrt((ex - sx) ** 2 + (ey - sy) ** 2)
        segs[i] = [sx, sy, ex, ey, dx, dy, l2, (1.0 / l2 if l2 > 0 else 0.0), cum, s.banking, chord, 0.0]
        seg64[i] = [sx, sy, ex, ey, s.banking, 0.0]
        cum += chord
    # --- uniform grid -------------------------------------------------------
    cs = _wall_corners(boxes)
    pad = 2.0
    x0 = float(cs[:, :, 0].min()) - pad
    y0 = float(cs[:, :, 1].min()) - pad
    x1 = float(cs[:, :, 0].max()) + pad
    y1 = float(cs[:, :, 1].max()) + pad
    nx = max(1, int(math.ceil((x1 - x0) / cell)))
    ny = max(1, int(math.ceil((y1 - y0) / cell)))
    lists: List[List[int]] = [[] for _ in range(nx * ny)]
    margin = 0.25
    for w in range(n):
        lx, ly, ux, uy = [float(v) for v in fat[w]]
        ix0 = max(0, int(math.floor((lx - margin - x0) / cell)))
        ix1 = min(nx - 1, int(math.floor((ux + margin - x0) / cell)))
        iy0 = max(0, int(math.floor((ly - margin - y0) / cell)))
        iy1 = min(ny - 1, int(math.floor((uy + margin - y0) / cell)))
        for iy in range(iy0, iy1 + 1):
            for ix in range(ix0, ix1 + 1):
                lo = (x0 + ix * cell, y0 + iy * cell)
                hi = (lo[0] + cell, lo[1] + cell)
                if _obb_overlaps_cell(boxes[w], lo, hi, margin):
                    lists[iy * nx + ix].append(w)
    starts = np.zeros(nx * ny + 1, dtype=np.int64)
    for i, l in enumerate(lists):
        starts[i + 1] = starts[i] + len(l)
    if starts[-1] >= 65536:
        raise ValueError("grid item list exceeds 16-bit offsets; use a larger cell")
    items = np.asarray([w for l in lists for w in l], dtype=np.uint16)
    cell_start = starts.astype(np.uint16)

    # --- blob ---------------------------------------------------------------
    def pad4(nwords):
        return (nwords + 3) // 4 * 4

    off_segs = HDR_WORDS
    off_seg64 = off_segs + MAX_SEGS * SEG_STRIDE               # even word offset => 8-byte aligned doubles
    off_walls = off_seg64 + pad4(MAX_SEGS * SEG64_STRIDE * 2)
    # device layout of the grid: one u32 per cell = first block | n_blocks << 16, items in blocks of ITEM_BLOCK u16
    blk_first = np.zeros(nx * ny, dtype=np.int64)
    blk_count = np.zeros(nx * ny, dtype=np.int64)
    padded: List[int] = []
    for i, l in enumerate(lists):
        nb = (len(l) + ITEM_BLOCK - 1) // ITEM_BLOCK
        blk_first[i], blk_count[i] = len(padded) // ITEM_BLOCK, nb
        padded.extend(l)
        if l:                                        # pad the last block by repeating its first wall (a harmless re-test)
            padded.extend([l[(nb - 1) * ITEM_BLOCK]] * (nb * ITEM_BLOCK - len(l)))
    if len(padded) // ITEM_BLOCK >= 65536:
        raise ValueError("grid item list exceeds 16-bit block offsets; use a larger cell")
    off_cells = off_walls + pad4(n * WALL_STRIDE)
    cells_words = pad4(nx * ny)
    off_items = off_cells + cells_words
    items_words = pad4((len(padded) + 1) // 2)
    # per cell: the segments whose chord can be nearest to some point of the cell (a conservative superset: distance from
    # the cell centre within two half-diagonals + 0.5 m of the smallest), so the per-step nearest-segment search tests
    # 1-3 chords instead of all of them and still finds the same first strict minimum and the same runner-up band
    segmask = np.zeros(nx * ny, dtype=np.uint16)
    hd = cell * math.sqrt(2.0) / 2.0
    for iy in range(ny):
        for ix in range(nx):
            pcx, pcy = x0 + (ix + 0.5) * cell, y0 + (iy + 0.5) * cell
            ds = []
            for sg in segs_src:
                sx, sy = sg.start
                dx, dy = sg.end[0] - sx, sg.end[1] - sy
                l2 = dx * dx + dy * dy
                t = 0.0 if l2 < 1e-6 else max(0.0, min(1.0, ((pcx - sx) * dx + (pcy - sy) * dy) / l2))
                ds.append(math.hypot(pcx - (sx + t * dx), pcy - (sy + t * dy)))
            lim = min(ds) + 2.0 * hd + 0.5
            m = 0
            for i, d in enumerate(ds):
                if d <= lim:
                    m |= 1 << i
            segmask[iy * nx + ix] = m
    off_segmask = off_items + items_words
    segmask_words = pad4((nx * ny + 1) // 2)
    off_aabb = off_segmask + segmask_words      # wall fat AABBs (broad phase)
    total = off_aabb + pad4(n * 4)
    blob = np.zeros(total, dtype=np.float32)
    hi = blob.view(np.int32)
    hi[H_NWALLS] = n
    hi[H_NSEGS] = len(segs_src)
    hi[H_GNX] = nx
    hi[H_GNY] = ny
    hi[H_HASBANK] = 1 if track.has_banking else 0
    hi[H_WORDS] = total
    hi[H_OFF_SEGS] = off_segs
    hi[H_OFF_WALLS] = off_walls
    hi[H_OFF_AABB] = off_aabb
    hi[H_OFF_CELLS] = off_cells
    hi[H_OFF_ITEMS] = off_items
    hi[H_NITEMS] = len(padded)
    hi[H_STAGE_WORDS] = total
    hi[H_OFF_SEG64] = off_seg64
    hi[H_OFF_SEGMASK] = off_segmask
    blob[H_GX0] = x0
    blob[H_GY0] = y0
    blob[H_INVCELL] = 1.0 / cell
    blob[H_CELL] = cell
    blob[H_LTOT] = track.total_length
    blob[H_MINLAP] = track.total_length * K.MIN_LAP_DISTANCE_FRACTION
    blob[H_HALF_LTOT] = track.total_length / 2
    sl = track.startline
    if sl is not None:
        dx, dy = sl.end[0] - sl.start[0], sl.end[1] - sl.start[1]
        blob[H_SLX0], blob[H_SLY0], blob[H_SLDX], blob[H_SLDY] = sl.start[0], sl.start[1], dx, dy
        blob[H_SLLEN2] = dx * dx + dy * dy
        blob[H_SLHALFW] = sl.width / 2.0
    else:
        blob[H_SLHALFW] = -1.0
    blob[off_segs:off_segs + MAX_SEGS * SEG_STRIDE] = segs.reshape(-1)
    s64 = np.zeros((MAX_SEGS, SEG64_STRIDE), dtype=np.float64)
    cum64 = 0.0
    for i, sg in enumerate(segs_src):
        s64[i] = [sg.start[0], sg.start[1], sg.end[0], sg.end[1], cum64]
        cum64 += math.sqrt((sg.end[0] - sg.start[0]) ** 2 + (sg.end[1] - sg.start[1]) ** 2)     # car_env.py:1594-1600
    blob[off_seg64:off_seg64 + MAX_SEGS * SEG64_STRIDE * 2] = s64.reshape(-1).view(np.float32)
    wrows = np.zeros((n, WALL_STRIDE), dtype=np.float32)
    wrows[:, :7] = boxes
    # the half-length the sensor rays use: 0.1 mm longer, so that collinear boxes abutting end to end (every straight) are
    # watertight for a ray that runs exactly in the plane of the joint (every car starts on one: x = 0, heading 0)
    wrows[:, 7] = boxes[:, 4]
    wrows[:, 4] = (boxes[:, 4].astype(np.float64) + 1e-4).astype(np.float32)
    blob[off_walls:off_walls + n * WALL_STRIDE] = wrows.reshape(-1)
    blob[off_aabb:off_aabb + n * 4] = fat.reshape(-1)
    blob.view(np.uint32)[off_cells: off_cells + nx * ny] = (blk_first | (blk_count << 16)).astype(np.uint32)
    it16 = np.zeros(items_words * 2, dtype=np.uint16)
    it16[:len(padded)] = padded
    blob.view(np.uint16)[off_items * 2: off_items * 2 + len(it16)] = it16
    blob.view(np.uint16)[off_segmask * 2: off_segmask * 2 + nx * ny] = segmask
    return TrackTable(track, boxes, fat, segs, seg64, (x0, y0), (nx, ny), cell_start, items, blob)


# ----------------------------------------------------------------------------
# Grid cell size: picked per track from a cost model of the ray loop
# ----------------------------------------------------------------------------
CELL_CANDIDATES = (12.0, 16.0, 24.0, 32.0, 48.0, 64.0)
MAX_TABLE_BYTES = 64 * 1024          # a CTA stages the whole table next to ~32 KB of its own state
# Three CTAs are resident per SM when a CTA's shared memory is at most 228 KB / 3 - 1 KB reserved - ~1 KB static, i.e. when
# the table is at most ~44.8 KB next to the 30.9 KB of csrc/ncg_b200.cu::smem_layout.  A batch that mixes tracks sizes
# every CTA for the largest table, so one track over this line costs all of them the third resident CTA.
THREE_CTA_TABLE_BYTES = 44800
THREE_CTA_COST_SLACK = 1.10          # a table that fits may cost this much more in the ray loop than the best one


def _sample_rays(tab: TrackTable, n_poses: int = 120, seed: int = 0):
    """Ray origins near the racing surface with random headings, 16 rays each, and their true hit distances."""
    rng = np.random.default_rng(seed)
    P = []
    for _ in range(n_poses):
        s = tab.seg64[rng.integers(0, len(tab.seg64))]
        u = rng.uniform()
        x = s[0] + u * (s[2] - s[0]) + rng.uniform(-4, 4)
        y = s[1] + u * (s[3] - s[1]) + rng.uniform(-4, 4)
        a = rng.uniform(-math.pi, math.pi)
        for k in range(K.NUM_SENSORS):
            P.append((x, y, math.cos(a - k * math.pi / 8), math.sin(a - k * math.pi / 8)))
    px, py, dx, dy = np.array(P).T
    b = tab.boxes.astype(np.float64)
    c, s_ = b[:, 2][None, :], b[:, 3][None, :]
    ax, ay = b[:, 0][None, :] - px[:, None], b[:, 1][None, :] - py[:, None]
    mx, my = c * ax + s_ * ay, c * ay - s_ * ax
    ex, ey = c * dx[:, None] + s_ * dy[:, None] + 1e-30, c * dy[:, None] - s_ * dx[:, None] + 1e-30
    x0, x1, y0, y1 = (mx - b[:, 4]) / ex, (mx + b[:, 4]) / ex, (my - b[:, 5]) / ey, (my + b[:, 5]) / ey
    tn = np.maximum(np.minimum(x0, x1), np.minimum(y0, y1))
    tf = np.minimum(np.maximum(x0, x1), np.maximum(y0, y1))
    hit = np.minimum(np.where((tn > 0) & (tn <= tf), tn, np.inf).min(axis=1), K.SENSOR_MAX_DISTANCE)
    return px, py, dx, dy, hit


def ray_loop_cost(tab: TrackTable, cell: float, rays) -> float:
    """Modelled cost of csrc/ncg_car.cuh::cast_rays on this table: iterations of its flattened loop (one per block of four
    walls, at least one per visited cell; same traversal rule as the device code) for the busiest lane of a car, averaged
    over the sampled cars.  A warp advances at the pace of its busiest lane, so the long rays are what matter; a lane
    owns rays q and q+4 (the kernel's 2-rays-per-lane mapping)."""
    px, py, dx, dy, hit = rays
    x0, y0 = tab.grid_origin
    nx, ny = tab.grid_dims
    nblk = (np.diff(tab.cell_start.astype(np.int64)) + ITEM_BLOCK - 1) // ITEM_BLOCK
    iters = np.zeros(len(px))
    for i in range(len(px)):
        gx, gy = (px[i] - x0) / cell, (py[i] - y0) / cell
        ix, iy = int(math.floor(gx)), int(math.floor(gy))
        if not (0 <= ix < nx and 0 <= iy < ny):
            continue
        fx, fy = gx - ix, gy - iy
        tdx = cell / abs(dx[i]) if dx[i] != 0 else math.inf
        tdy = cell / abs(dy[i]) if dy[i] != 0 else math.inf
        tmx = ((1 - fx) if dx[i] > 0 else fx) * tdx if dx[i] != 0 else math.inf
        tmy = ((1 - fy) if dy[i] > 0 else fy) * tdy if dy[i] != 0 else math.inf
        sx, sy = (1 if dx[i] > 0 else -1), (1 if dy[i] > 0 else -1)
        n = 0
        while True:
            n += max(1, int(nblk[iy * nx + ix]))
            texit = min(tmx, tmy)
            if hit[i] <= texit or texit >= K.SENSOR_MAX_DISTANCE:
                break
            if tmx < tmy:
                ix += sx
                tmx += tdx
            else:
                iy += sy
                tmy += tdy
            if not (0 <= ix < nx and 0 <= iy < ny):
                break
        iters[i] = n
    it = iters.reshape(-1, K.NUM_SENSORS)
    lanes = [q if q < 4 else q + 4 for q in range(8)]
    per_lane = it[:, lanes] + it[:, [q + 4 for q in lanes]]
    return float(per_lane.max(axis=1).mean())


def build_best_track_table(track: Track) -> TrackTable:
    """build_track_table at the candidate cell size with the lowest modelled ray cost that fits the staging budget
    (the superspeedways end up at 48 m, the half-mile tracks with their ~1 m chords at 16-24 m)."""
    forced = os.environ.get("NCG_GRID_CELL")
    if forced:
        return build_track_table(track, cell=float(forced))
    best, small, rays = None, None, None
    for cell in CELL_CANDIDATES:
        try:
            tab = build_track_table(track, cell=cell)
        except ValueError:
            continue
        if tab.blob.nbytes > MAX_TABLE_BYTES:
            continue
        if rays is None:
            rays = _sample_rays(tab)
        cost = ray_loop_cost(tab, cell, rays)
        if best is None or cost < best[0]:
            best = (cost, tab)
        if tab.blob.nbytes <= THREE_CTA_TABLE_BYTES and (small is None or cost < small[0]):
            small = (cost, tab)
    if best is None:
        return build_track_table(track)
    if small is not None and small[0] <= THREE_CTA_COST_SLACK * best[0]:
        return small[1]
    return best[1]


_TABLE_CACHE: Dict[str, TrackTable] = {}


def get_track_table(path_or_name: str) -> TrackTable:
    key = os.path.abspath(path_or_name) if os.path.exists(path_or_name) else track_name_of(path_or_name)
    if key not in _TABLE_CACHE:
        _TABLE_CACHE[key] = build_best_track_table(load_track(path_or_name))
    return _TABLE_CACHE[key]
