"""Emulates the GPU path's rounding points on the CPU to size the operand-precision error vs the fp32 oracle."""
import sys, numpy as np, torch, torch.nn.functional as F
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from oracle import unet as ounet, dice, geometry as ogeo

def rnd(t, dt):
    return t.to(dt).float() if dt is not None else t

def emu_forward(x, sd, dt, raw_dt=None):
    def cir(x, p, first=False):
        w = sd[p + '.weight']
        y = F.conv3d(x if first else rnd(x, dt), w if first else rnd(w, dt), None, padding=1)
        y = rnd(y, raw_dt)
        y = F.relu(F.instance_norm(y, eps=1e-5))
        return rnd(y, dt)
    def ct(x, p):
        return rnd(F.conv_transpose3d(rnd(x, dt), rnd(sd[p + '.weight'], dt), sd[p + '.bias'], stride=2), dt)
    with torch.no_grad():
        c1 = cir(x, 'double_conv1.convolution.0', first=True)
        c1 = cir(c1, 'double_conv1.convolution.3')
        c2 = cir(F.max_pool3d(c1, 2), 'double_conv2.convolution.0')
        c2 = cir(c2, 'double_conv2.convolution.3')
        b = cir(F.max_pool3d(c2, 2), 'bottom_layer.convolution.0')
        b = cir(b, 'bottom_layer.convolution.3')
        b = cir(b, 'bottom_layer.convolution.6')
        e2 = cir(torch.cat([c2, ct(b, 't_conv2')], 1), 'ex_double_conv2.convolution.0')
        e2 = cir(e2, 'ex_double_conv2.convolution.3')
        cat1 = torch.cat([c1, ct(e2, 't_conv1')], 1)
        # last conv: raw fp32 -> IN+ReLU fp32 -> head fp32
        y = F.conv3d(rnd(cat1, dt), rnd(sd['ex_conv1_1.convolution.0.weight'], dt), None, padding=1)
        y = F.relu(F.instance_norm(y, eps=1e-5))
        o = F.conv3d(y, sd['one_by_one.weight'], sd['one_by_one.bias'])
        o = F.conv3d(o, sd['one_by_one_2.weight'], sd['one_by_one_2.bias'])
        return torch.sigmoid(o)

def psnr(a, b):
    return 10 * np.log10(1.0 / float(((a - b) ** 2).mean()))

if __name__ == "__main__":
    case = sys.argv[1]
    sd = ounet.random_state_dict(seed=0, bias_std=0.1)
    if case == 'small':
        rng = np.random.default_rng(0); vol = (rng.random((40, 52, 30)) ** 3 * 65535).astype(np.uint16)
        g = ogeo.dice_geometry(vol.shape, 24, 6, 4); cubes = range(g.n_cubes)
    elif case == 'c1':
        rng = np.random.default_rng(4); vol = (rng.random((128, 128, 128)) ** 3 * 65535).astype(np.uint16)
        g = ogeo.dice_geometry(vol.shape, 120, 15, 10); cubes = [int(c) for c in sys.argv[2:]] or [5]
    for i in cubes:
        x = torch.from_numpy(dice.dice_cube_gather(vol, g, i))[None]
        ref = ounet.unet_deconv_forward(x, sd)
        res = []
        for name, dt in (('bf16', torch.bfloat16), ('fp16', torch.float16)):
            y = emu_forward(x, sd, dt)
            res.append('%s max %.4f psnr %.1f' % (name, float((y - ref).abs().max()), psnr(y, ref)))
        print('cube', i, 'zero-frac %.2f' % float((x == 0).float().mean()), ' | '.join(res), flush=True)
