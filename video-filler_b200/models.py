"""Network builders: the generator / discriminator of ``train.lua:87-202`` (image, inpaintCenter)
and of ``train_vid_weighted.lua:112-239`` / ``train_deepernet.lua`` (channel-stacked video clips)."""
import numpy as np

from . import nn


def default_opt(variant="image", **kw):
    """Option tables (train.lua:6-35, train_vid_weighted.lua:15-54) at the benchmark settings."""
    if variant == "image":
        opt = dict(variant="image", batchSize=64, fineSize=128, nBottleneck=4000, nef=64, ngf=64, ndf=64, nc=3,
                   predLen=1, wtl2=0.999, overlapPred=4, lr=0.0002, beta1=0.5, weight_nomask=0.05, wtgdl=0.0)
    else:
        opt = dict(variant="video", batchSize=64, fineSize=128, nBottleneck=4000, nef=64, ngf=64, ndf=64, nc=3,
                   predLen=4, wtl2=0.999, overlapPred=0, lr=0.0002, beta1=0.5, weight_nomask=0.05, wtgdl=0.0,
                   maskValue=110.0 / 255.0)
    opt.update(kw)
    return opt


def net_channels(opt):
    return opt["nc"] * opt["predLen"] if opt["variant"] == "video" else opt["nc"]


def build_netG(opt):
    nc, nef, ngf, nB = net_channels(opt), opt["nef"], opt["ngf"], opt["nBottleneck"]
    Conv, Full, BN = nn.SpatialConvolution, nn.SpatialFullConvolution, nn.SpatialBatchNormalization
    netE = nn.Sequential()
    netE.add(Conv(nc, nef, 4, 4, 2, 2, 1, 1)).add(nn.LeakyReLU(0.2, True))
    for cin, cout in ((nef, nef), (nef, nef * 2), (nef * 2, nef * 4), (nef * 4, nef * 8)):
        netE.add(Conv(cin, cout, 4, 4, 2, 2, 1, 1)).add(BN(cout)).add(nn.LeakyReLU(0.2, True))
    netE.add(Conv(nef * 8, nB, 4, 4))
    netG = nn.Sequential()
    nz_size = nB
    if opt.get("noiseGen"):               # train.lua:109-124
        nz = opt.get("nz", 100)
        netG_noise = nn.Sequential().add(Conv(nz, nz, 1, 1, 1, 1, 0, 0))
        netG.add(nn.ParallelTable().add(netE).add(netG_noise))
        netG.add(nn.JoinTable(2))
        nz_size = nB + nz
    else:
        netG.add(netE)
    netG.add(BN(nz_size)).add(nn.LeakyReLU(0.2, True))
    netG.add(Full(nz_size, ngf * 8, 4, 4)).add(BN(ngf * 8)).add(nn.ReLU(True))
    chain = [(ngf * 8, ngf * 4), (ngf * 4, ngf * 2), (ngf * 2, ngf)]
    if opt["variant"] == "video":
        chain.append((ngf, ngf))          # train_vid_weighted.lua:171-172
    for cin, cout in chain:
        netG.add(Full(cin, cout, 4, 4, 2, 2, 1, 1)).add(BN(cout)).add(nn.ReLU(True))
    netG.add(Full(ngf, nc, 4, 4, 2, 2, 1, 1))
    netG.add(nn.Tanh())
    return netG


def build_netD(opt):
    nc, ndf = net_channels(opt), opt["ndf"]
    Conv, BN = nn.SpatialConvolution, nn.SpatialBatchNormalization
    netD = nn.Sequential()
    if opt["variant"] == "video":
        mylayer = ndf // 2                # train_vid_weighted.lua:213-221
        netD.add(Conv(nc, mylayer, 4, 4, 2, 2, 1, 1)).add(nn.LeakyReLU(0.2, True))
        netD.add(Conv(mylayer, ndf, 4, 4, 2, 2, 1, 1)).add(nn.LeakyReLU(0.2, True))
    elif opt.get("conditionAdv"):         # train.lua:158-180
        netD_ctx = nn.Sequential().add(Conv(nc, ndf, 5, 5, 2, 2, 2, 2))
        netD_pred = nn.Sequential().add(Conv(nc, ndf, 5, 5, 2, 2, 2 + 32, 2 + 32))
        netD.add(nn.ParallelTable().add(netD_ctx).add(netD_pred))
        netD.add(nn.JoinTable(2))
        netD.add(nn.LeakyReLU(0.2, True))
        netD.add(Conv(ndf * 2, ndf, 4, 4, 2, 2, 1, 1)).add(BN(ndf)).add(nn.LeakyReLU(0.2, True))
    else:
        netD.add(Conv(nc, ndf, 4, 4, 2, 2, 1, 1)).add(nn.LeakyReLU(0.2, True))
    for cin, cout in ((ndf, ndf * 2), (ndf * 2, ndf * 4), (ndf * 4, ndf * 8)):
        netD.add(Conv(cin, cout, 4, 4, 2, 2, 1, 1)).add(BN(cout)).add(nn.LeakyReLU(0.2, True))
    netD.add(Conv(ndf * 8, 1, 4, 4))
    netD.add(nn.Sigmoid())
    netD.add(nn.View(1).setNumInputDims(3))
    return netD


def weights_init(net, rng):
    """train.lua:58-67 with a host numpy Generator (so runs are reproducible against any host-side copy)."""
    def init(m):
        name = m.type_name()
        if "Convolution" in name:
            m.weight.copy_(rng.normal(0.0, 0.02, m.weight.shape).astype(np.float32))
            m.bias.fill(0)
        elif "BatchNormalization" in name:
            if m.weight is not None:
                m.weight.copy_(rng.normal(1.0, 0.02, m.weight.shape).astype(np.float32))
            if m.bias is not None:
                m.bias.fill(0)
    net.apply(init)


def zero_conv_bias(net):
    """train.lua:279-280."""
    def z(m):
        if "Convolution" in m.type_name():
            m.bias.zero()
    net.apply(z)
