"""Network builders, weight init and synthetic inputs restating the reference scripts.

* image variant  : ``train.lua:87-202``  (inpaintCenter, 128x128 context -> 64x64 centre)
* video variant  : ``train_vid_weighted.lua:112-239`` / ``train_deepernet.lua`` (nc = 3*predLen
  channel-stacked clips, full-size output, extra decoder layer, 6-conv D)

Test infrastructure only -- see ``oracle/__init__.py``.
"""
import numpy as np

from . import nn


def default_opt(variant='image', **kw):
    """Option tables of train.lua:6-35 / train_vid_weighted.lua:15-54 (benchmark settings of SURVEY 8d)."""
    if variant == 'image':
        opt = dict(variant='image', batchSize=64, fineSize=128, nBottleneck=4000, nef=64, ngf=64, ndf=64, nc=3,
                   predLen=1, wtl2=0.999, overlapPred=4, lr=0.0002, beta1=0.5, weight_nomask=0.05, wtgdl=0.0)
    else:
        opt = dict(variant='video', batchSize=64, fineSize=128, nBottleneck=4000, nef=64, ngf=64, ndf=64, nc=3,
                   predLen=4, wtl2=0.999, overlapPred=0, lr=0.0002, beta1=0.5, weight_nomask=0.05, wtgdl=0.0,
                   maskValue=110.0 / 255.0)
    opt.update(kw)
    return opt


def build_netG(opt, dtype=np.float32):
    """train.lua:87-150 (image) / train_vid_weighted.lua:112-176 (video; one more decoder stage)."""
    nc = opt['nc'] * opt['predLen'] if opt['variant'] == 'video' else opt['nc']
    nef, ngf, nB = opt['nef'], opt['ngf'], opt['nBottleneck']
    C, FC, BN = nn.SpatialConvolution, nn.SpatialFullConvolution, nn.SpatialBatchNormalization
    netE = nn.Sequential()
    netE.add(C(nc, nef, 4, 4, 2, 2, 1, 1, dtype=dtype)).add(nn.LeakyReLU(0.2, True))
    netE.add(C(nef, nef, 4, 4, 2, 2, 1, 1, dtype=dtype)).add(BN(nef, dtype=dtype)).add(nn.LeakyReLU(0.2, True))
    netE.add(C(nef, nef * 2, 4, 4, 2, 2, 1, 1, dtype=dtype)).add(BN(nef * 2, dtype=dtype)).add(nn.LeakyReLU(0.2, True))
    netE.add(C(nef * 2, nef * 4, 4, 4, 2, 2, 1, 1, dtype=dtype)).add(BN(nef * 4, dtype=dtype)).add(nn.LeakyReLU(0.2, True))
    netE.add(C(nef * 4, nef * 8, 4, 4, 2, 2, 1, 1, dtype=dtype)).add(BN(nef * 8, dtype=dtype)).add(nn.LeakyReLU(0.2, True))
    netE.add(C(nef * 8, nB, 4, 4, dtype=dtype))
    netG = nn.Sequential()
    nz_size = nB
    if opt.get('noiseGen'):
        # train.lua:109-124: a 1x1 conv on the noise vector runs beside the encoder, joined along the channel axis
        nz = opt.get('nz', 100)
        netG_noise = nn.Sequential().add(C(nz, nz, 1, 1, 1, 1, 0, 0, dtype=dtype))
        netG.add(nn.ParallelTable().add(netE).add(netG_noise))
        netG.add(nn.JoinTable(2))
        nz_size = nB + nz
    else:
        netG.add(netE)
    netG.add(BN(nz_size, dtype=dtype)).add(nn.LeakyReLU(0.2, True))
    netG.add(FC(nz_size, ngf * 8, 4, 4, dtype=dtype)).add(BN(ngf * 8, dtype=dtype)).add(nn.ReLU(True))
    netG.add(FC(ngf * 8, ngf * 4, 4, 4, 2, 2, 1, 1, dtype=dtype)).add(BN(ngf * 4, dtype=dtype)).add(nn.ReLU(True))
    netG.add(FC(ngf * 4, ngf * 2, 4, 4, 2, 2, 1, 1, dtype=dtype)).add(BN(ngf * 2, dtype=dtype)).add(nn.ReLU(True))
    netG.add(FC(ngf * 2, ngf, 4, 4, 2, 2, 1, 1, dtype=dtype)).add(BN(ngf, dtype=dtype)).add(nn.ReLU(True))
    if opt['variant'] == 'video':
        netG.add(FC(ngf, ngf, 4, 4, 2, 2, 1, 1, dtype=dtype)).add(BN(ngf, dtype=dtype)).add(nn.ReLU(True))
    netG.add(FC(ngf, nc, 4, 4, 2, 2, 1, 1, dtype=dtype))
    netG.add(nn.Tanh())
    return netG


def build_netD(opt, dtype=np.float32):
    """train.lua:157-202 (image: 64x64 input) / train_vid_weighted.lua:183-239 (video: 128x128 input)."""
    nc = opt['nc'] * opt['predLen'] if opt['variant'] == 'video' else opt['nc']
    ndf = opt['ndf']
    C, BN = nn.SpatialConvolution, nn.SpatialBatchNormalization
    netD = nn.Sequential()
    if opt['variant'] == 'video':
        mylayer = ndf // 2
        netD.add(C(nc, mylayer, 4, 4, 2, 2, 1, 1, dtype=dtype)).add(nn.LeakyReLU(0.2, True))
        netD.add(C(mylayer, ndf, 4, 4, 2, 2, 1, 1, dtype=dtype)).add(nn.LeakyReLU(0.2, True))
    elif opt.get('conditionAdv'):
        # train.lua:158-180: D also sees the context; the 64x64 prediction is padded by 32 so that both branches give 64x64 maps
        netD_ctx = nn.Sequential().add(C(nc, ndf, 5, 5, 2, 2, 2, 2, dtype=dtype))
        netD_pred = nn.Sequential().add(C(nc, ndf, 5, 5, 2, 2, 2 + 32, 2 + 32, dtype=dtype))
        netD.add(nn.ParallelTable().add(netD_ctx).add(netD_pred))
        netD.add(nn.JoinTable(2))
        netD.add(nn.LeakyReLU(0.2, True))
        netD.add(C(ndf * 2, ndf, 4, 4, 2, 2, 1, 1, dtype=dtype)).add(BN(ndf, dtype=dtype)).add(nn.LeakyReLU(0.2, True))
    else:
        netD.add(C(nc, ndf, 4, 4, 2, 2, 1, 1, dtype=dtype)).add(nn.LeakyReLU(0.2, True))
    netD.add(C(ndf, ndf * 2, 4, 4, 2, 2, 1, 1, dtype=dtype)).add(BN(ndf * 2, dtype=dtype)).add(nn.LeakyReLU(0.2, True))
    netD.add(C(ndf * 2, ndf * 4, 4, 4, 2, 2, 1, 1, dtype=dtype)).add(BN(ndf * 4, dtype=dtype)).add(nn.LeakyReLU(0.2, True))
    netD.add(C(ndf * 4, ndf * 8, 4, 4, 2, 2, 1, 1, dtype=dtype)).add(BN(ndf * 8, dtype=dtype)).add(nn.LeakyReLU(0.2, True))
    netD.add(C(ndf * 8, 1, 4, 4, dtype=dtype))
    netD.add(nn.Sigmoid())
    netD.add(nn.View(1).setNumInputDims(3))
    return netD


def weights_init(net, rng):
    """train.lua:58-67: conv/full-conv weight N(0,0.02), bias 0; BN gamma N(1,0.02), beta 0."""
    def init(m):
        name = m.type_name()
        if 'Convolution' in name:
            m.weight[...] = rng.normal(0.0, 0.02, m.weight.shape).astype(m.weight.dtype)
            m.bias[...] = 0
        elif 'BatchNormalization' in name:
            if m.weight is not None:
                m.weight[...] = rng.normal(1.0, 0.02, m.weight.shape).astype(m.weight.dtype)
            if m.bias is not None:
                m.bias[...] = 0
    net.apply(init)


def zero_conv_bias(net):
    """train.lua:279-280: every Convolution bias is zeroed at the start of both closures."""
    def z(m):
        if 'Convolution' in m.type_name():
            m.bias[...] = 0
    net.apply(z)


MEAN_FILL = (2 * 117.0 / 255.0 - 1.0, 2 * 104.0 / 255.0 - 1.0, 2 * 123.0 / 255.0 - 1.0)


def synth_image_batch(B, fineSize, overlapPred, rng, dtype=np.float32):
    """SURVEY 8d cfg1/2 + train.lua:286-290: U(-1,1) images, centre cloned, inner centre mean-filled."""
    real = rng.uniform(-1.0, 1.0, (B, 3, fineSize, fineSize)).astype(dtype)
    q, h = fineSize // 4, fineSize // 2
    real_center = real[:, :, q:q + h, q:q + h].copy()
    real_ctx = real.copy()
    for c in range(3):
        real_ctx[:, c, q + overlapPred:q + h - overlapPred, q + overlapPred:q + h - overlapPred] = MEAN_FILL[c]
    return real_ctx, real_center


def random_block_mask(fineSize, rng):
    """datavid/donkey_folder.lua:114-129 rule: 2-10 square blocks of floor(fineSize/6) px."""
    m = np.zeros((fineSize, fineSize), np.uint8)
    blk = fineSize // 6
    for _ in range(int(rng.integers(2, 11))):
        y = int(rng.integers(0, fineSize - blk + 1))
        x = int(rng.integers(0, fineSize - blk + 1))
        m[y:y + blk, x:x + blk] = 1
    return m


def synth_video_batch(B, nc, fineSize, maskValue, rng, dtype=np.float32):
    """SURVEY 8d cfg3 + datavid/dataset.lua:426 contract: (masked, full, mask uint8), mask shared over channels."""
    full = rng.uniform(-1.0, 1.0, (B, nc, fineSize, fineSize)).astype(dtype)
    mask = np.empty((B, nc, fineSize, fineSize), np.uint8)
    for b in range(B):
        mask[b] = random_block_mask(fineSize, rng)[None]
    fill = np.asarray(2 * maskValue - 1, dtype)
    masked = np.where(mask != 0, fill, full).astype(dtype)
    return masked, full, mask
