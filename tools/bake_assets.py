#!/usr/bin/env python3
"""Bake the reference's input assets into the compact forms this repo ships under assets/.

The GPU box has no /root/reference, so the scene inputs travel with the repo:
  *.obj  -> assets/<name>.mesh   binary: magic 'PTM1', u32 n_pos, n_idx, n_tex, n_nrm, then
                                 f32 positions[3*n_pos], u32 indices[n_idx], f32 texcoords[2*n_tex],
                                 f32 normals[3*n_nrm]  (what tobj 4.0.2 yields with
                                 OFFLINE_RENDERING_LOAD_OPTIONS: triangulated, f32, position indices)
  images -> assets/<name>.png    RGB8 exactly as image 0.25.5's DynamicImage::to_rgb8() yields
                                 (reference src/texture.rs:62-69): .hdr is clamped to [0,1] and
                                 quantised round(x*255) (Q22); RGBA drops alpha.
  envmap.jpg (7616x3808) stays a JPEG (a lossless PNG would be ~40 MB); decoded at load time.

Usage: python tools/bake_assets.py [--src /root/reference/assets] [--raw]
  --raw additionally writes assets/baked/*.rgb8 ('PTI1', u32 w, u32 h, raw RGB) for the C++ CLI.
"""
import argparse, os, shutil, struct, sys
import numpy as np

def parse_obj(path):
    pos, tex, nrm, idx = [], [], [], []
    for line in open(path):
        p = line.split()
        if not p: continue
        if p[0] == 'v': pos.append([np.float32(x) for x in p[1:4]])
        elif p[0] == 'vt': tex.append([np.float32(x) for x in p[1:3]])
        elif p[0] == 'vn': nrm.append([np.float32(x) for x in p[1:4]])
        elif p[0] == 'f':
            vs = []
            for tok in p[1:]:
                i = int(tok.split('/')[0])
                vs.append(i - 1 if i > 0 else len(pos) + i)
            for k in range(1, len(vs) - 1):  # fan triangulation, as tobj does
                idx += [vs[0], vs[k], vs[k + 1]]
    return (np.array(pos, np.float32).reshape(-1, 3), np.array(idx, np.uint32),
            np.array(tex, np.float32).reshape(-1, 2), np.array(nrm, np.float32).reshape(-1, 3))

def write_mesh(path, pos, idx, tex, nrm):
    with open(path, 'wb') as f:
        f.write(b'PTM1' + struct.pack('<4I', len(pos), len(idx), len(tex), len(nrm)))
        f.write(pos.tobytes()); f.write(idx.tobytes()); f.write(tex.tobytes()); f.write(nrm.tobytes())

def to_rgb8(path):
    """image::ImageReader::open(path).decode().to_rgb8()"""
    if path.endswith('.hdr'):
        import cv2
        img = cv2.imread(path, cv2.IMREAD_UNCHANGED)[:, :, ::-1]  # BGR f32 -> RGB
        return np.round(np.clip(img, 0.0, 1.0) * 255.0).astype(np.uint8)
    from PIL import Image
    Image.MAX_IMAGE_PIXELS = None
    return np.asarray(Image.open(path).convert('RGB'), dtype=np.uint8)

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--src', default='/root/reference/assets')
    ap.add_argument('--dst', default=os.path.join(os.path.dirname(__file__), '..', 'assets'))
    ap.add_argument('--raw', action='store_true')
    a = ap.parse_args()
    from PIL import Image
    os.makedirs(a.dst, exist_ok=True)
    if os.path.isdir(a.src):
        for name in ['bunny', 'spot', 'cow', 'teapot']:
            pos, idx, tex, nrm = parse_obj(os.path.join(a.src, name + '.obj'))
            write_mesh(os.path.join(a.dst, name + '.mesh'), pos, idx, tex, nrm)
            print(f'{name}: {len(pos)} v, {len(idx)//3} f, {len(tex)} vt, {len(nrm)} vn')
        for src, dst in [('earthmap.jpg', 'earthmap.png'), ('grace_probe_latlong.hdr', 'grace_probe_latlong.png'),
                         ('bricks/color.png', 'bricks_color.png'), ('bricks/normal.png', 'bricks_normal.png')]:
            rgb = to_rgb8(os.path.join(a.src, src))
            Image.fromarray(rgb).save(os.path.join(a.dst, dst), optimize=True)
            print(f'{dst}: {rgb.shape[1]}x{rgb.shape[0]}')
        shutil.copyfile(os.path.join(a.src, 'envmap.jpg'), os.path.join(a.dst, 'envmap.jpg'))
    else:
        print(f'{a.src} not present; using the already-baked files in {a.dst}', file=sys.stderr)
    if a.raw:
        os.makedirs(os.path.join(a.dst, 'baked'), exist_ok=True)
        for name in ['earthmap.png', 'grace_probe_latlong.png', 'bricks_color.png', 'bricks_normal.png', 'envmap.jpg']:
            rgb = to_rgb8(os.path.join(a.dst, name))
            out = os.path.join(a.dst, 'baked', os.path.splitext(name)[0] + '.rgb8')
            with open(out, 'wb') as f:
                f.write(b'PTI1' + struct.pack('<2I', rgb.shape[1], rgb.shape[0])); f.write(rgb.tobytes())
            print('raw', out)

if __name__ == '__main__':
    main()
