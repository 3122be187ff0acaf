"""The G+D training step restated: ``optim.adam(fDx,...)`` + ``optim.adam(fGx,...)``.

Follows ``train.lua:278-410,421-424`` (image variant) and
``train_vid_weighted.lua:373-537`` (video variant), including the step-level
invariants of SURVEY.md 9.9 (bias zeroing, D-gradient accumulation over the real
and fake pass, reuse of D's fake-pass activations by fGx, GDL-loss-with-MSE-gradient).
Test infrastructure only -- see ``oracle/__init__.py``.
"""
import numpy as np

from . import nets, nn, ops


class StepOracle:
    def __init__(self, opt, seed=1234, dtype=np.float32):
        self.opt = opt
        self.dtype = dtype
        rng = np.random.default_rng(seed)
        self.netG = nets.build_netG(opt, dtype)
        self.netD = nets.build_netD(opt, dtype)
        nets.weights_init(self.netG, rng)
        nets.weights_init(self.netD, rng)
        self.criterion = nn.BCECriterion()
        self.criterionMSE = nn.MSECriterion()
        self.criterionGDL = nn.GDLCriterion(1)
        wtl2 = opt['wtl2']
        # train.lua:219-226
        self.lrG = opt['lr'] * 10 if 0 < wtl2 < 1 else opt['lr']
        self.lrD = opt['lr']
        self.stateG, self.stateD = {}, {}
        self.pD, self.gD = self.netD.getParameters()
        self.pG, self.gG = self.netG.getParameters()
        self.errD = self.errG = self.errG_l2 = self.errG_gdl = None
        self.errD_real = self.errD_fake = None

    # -- closures --------------------------------------------------------
    def _label(self, out, value):
        """The label tensor BCE receives.  At fineSize 128 the discriminator emits one value per sample and this is the scripts'
        `label` of batchSize entries (train.lua:236,302,336).  cfg4 (train_deepernet at 256 x 256, SURVEY 8a caveat): the 4x4 head
        sees an 8x8 map, emits [B,1,5,5], and nn.View(1):setNumInputDims(3) turns that into [25 B, 1] -- BCE's nElement check
        would reject a label of B entries, so the script cannot run as shipped.  CONVENTION adopted here and in the executor: every
        one of the 25 patch outputs of a sample carries that sample's label (label tensor of 25 B entries, sizeAverage over
        25 B), i.e. the patch-discriminator reading of the same net."""
        return np.full(out.size, value, self.dtype)

    def _d_in(self, center):
        return [self.input_ctx, center] if self.opt.get('conditionAdv') else center      # train.lua:300-311

    def _g_in(self):
        return [self.input_ctx, self.noise] if self.opt.get('noiseGen') else self.input_ctx   # train.lua:325-329

    def fDx_image(self, real_ctx, real_center, noise=None):
        o = self.opt
        if o.get('noiseGen'):
            assert noise is not None, "noiseGen: the step takes the noise tensor [B, nz, 1, 1] drawn at train.lua:319-323"
            self.noise = nn._q(noise.astype(self.dtype))
        nets.zero_conv_bias(self.netD)
        nets.zero_conv_bias(self.netG)
        self.gD[...] = 0
        B = real_ctx.shape[0]
        self.input_ctx = nn._q(real_ctx.astype(self.dtype))
        self.input_real_center = nn._q(real_center.astype(self.dtype))
        out = self.netD.forward(self._d_in(self.input_real_center))
        label = self._label(out, 1.0)
        self.errD_real = self.criterion.forward(out, label)
        df_do = self.criterion.backward(out, label)
        self.netD.backward(self._d_in(self.input_real_center), df_do)
        fake = self.netG.forward(self._g_in())
        self.input_center = fake.copy()
        out = self.netD.forward(self._d_in(self.input_center))
        label = self._label(out, 0.0)
        self.errD_fake = self.criterion.forward(out, label)
        df_do = self.criterion.backward(out, label)
        self.netD.backward(self._d_in(self.input_center), df_do)
        self.errD = self.errD_real + self.errD_fake
        return self.errD

    def fGx_image(self):
        o = self.opt
        nets.zero_conv_bias(self.netD)
        nets.zero_conv_bias(self.netG)
        self.gG[...] = 0
        B = self.input_ctx.shape[0]
        out = self.netD.output
        label = self._label(out, 1.0)
        self.errG = self.criterion.forward(out, label)
        df_do = self.criterion.backward(out, label)
        df_dg = self.netD.updateGradInput(self._d_in(self.input_center), df_do)
        if o.get('conditionAdv'):
            df_dg = df_dg[1]            # train.lua:371: df_dg[2], the prediction branch
        total = self.errG
        wtl2 = o['wtl2']
        if wtl2 != 0:
            self.errG_l2 = self.criterionMSE.forward(self.input_center, self.input_real_center)
            df_dg = ops.blend_l2_overlap(df_dg, self.input_center, self.input_real_center, wtl2, o['overlapPred'])
            total = ((1 - wtl2) * self.errG + wtl2 * self.errG_l2) if 0 < wtl2 < 1 else self.errG + wtl2 * self.errG_l2
        self.df_dg = df_dg = nn._q(df_dg)
        self.netG.backward(self._g_in(), df_dg)
        return total

    def fDx_video(self, real_ctx, real_full, real_mask):
        o = self.opt
        nets.zero_conv_bias(self.netD)
        nets.zero_conv_bias(self.netG)
        self.gD[...] = 0
        B = real_ctx.shape[0]
        self.input_ctx = nn._q(real_ctx.astype(self.dtype))
        self.input_real = nn._q(real_full.astype(self.dtype))
        self.input_mask = real_mask.astype(self.dtype)
        out = self.netD.forward(self.input_real)
        label = self._label(out, 1.0)
        self.errD_real = self.criterion.forward(out, label)
        df_do = self.criterion.backward(out, label)
        self.netD.backward(self.input_real, df_do)
        fake = self.netG.forward(self.input_ctx)
        if o['weight_nomask'] == 0:
            # train_vid_weighted.lua:429-432: composite fake into ground truth under the mask
            self.input_inpainted = ops.mask_composite(self.input_real, self.input_mask, fake)
        else:
            self.input_inpainted = fake.copy()
        out = self.netD.forward(self.input_inpainted)
        label = self._label(out, 0.0)
        self.errD_fake = self.criterion.forward(out, label)
        df_do = self.criterion.backward(out, label)
        self.netD.backward(self.input_inpainted, df_do)
        self.errD = self.errD_real + self.errD_fake
        return self.errD

    def fGx_video(self):
        o = self.opt
        nets.zero_conv_bias(self.netD)
        nets.zero_conv_bias(self.netG)
        self.gG[...] = 0
        B = self.input_ctx.shape[0]
        out = self.netD.output
        label = self._label(out, 1.0)
        self.errG = self.criterion.forward(out, label)
        df_do = self.criterion.backward(out, label)
        # train_vid_weighted.lua:481 passes input_real; the first module is a conv so values are unused
        df_dg = self.netD.updateGradInput(self.input_real, df_do)
        total = self.errG
        wtl2 = o['wtl2']
        assert o['overlapPred'] == 0, "video scripts require overlapPred == 0 (train_vid_weighted.lua:509)"
        if wtl2 != 0:
            self.errG_l2 = self.criterionMSE.forward(self.input_inpainted, self.input_real)
            df_dg, weights = ops.blend_l2_masked(df_dg, self.input_inpainted, self.input_real, self.input_mask,
                                                 wtl2, o['weight_nomask'])
            if weights is not None:
                self.input_mask = weights  # in place on input_mask (:494)
            total = ((1 - wtl2) * self.errG + wtl2 * self.errG_l2) if 0 < wtl2 < 1 else self.errG + wtl2 * self.errG_l2
        if o['wtgdl'] != 0:
            # :523-528 -- loss from GDL, gradient from criterionMSE:backward (reference quirk)
            self.errG_gdl = self.criterionGDL.forward(self.input_inpainted, self.input_real)
            df_dg_gdl = self.criterionMSE.backward(self.input_inpainted, self.input_real)
            total = total + o['wtgdl'] * self.errG_gdl
            df_dg = (df_dg + np.asarray(o['wtgdl'], self.dtype) * df_dg_gdl).astype(self.dtype)
        self.df_dg = df_dg = nn._q(df_dg)
        self.netG.backward(self.input_ctx, df_dg)
        return total

    # -- one full step (train.lua:421-424) -----------------------------------
    def step(self, *batch):
        o = self.opt
        if o['variant'] == 'image':
            self.fDx_image(*batch)
        else:
            self.fDx_video(*batch)
        ops.adam_step(self.pD, self.gD, self.stateD, self.lrD, o['beta1'])
        if o['variant'] == 'image':
            self.errG_total = self.fGx_image()
        else:
            self.errG_total = self.fGx_video()
        ops.adam_step(self.pG, self.gG, self.stateG, self.lrG, o['beta1'])
        return dict(errD=self.errD, errG=self.errG, errG_l2=self.errG_l2, errG_gdl=self.errG_gdl,
                    errD_real=self.errD_real, errD_fake=self.errD_fake, errG_total=self.errG_total)

    def synth_batch(self, rng, B=None):
        o = self.opt
        B = B or o['batchSize']
        if o['variant'] == 'image':
            return nets.synth_image_batch(B, o['fineSize'], o['overlapPred'], rng, self.dtype)
        return nets.synth_video_batch(B, o['nc'] * o['predLen'], o['fineSize'], o['maskValue'], rng, self.dtype)
