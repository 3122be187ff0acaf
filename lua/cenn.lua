--[[ cenn.lua -- LuaJIT-FFI binding of libcenn.so for the UNCHANGED reference scripts
     (train.lua, train_vid_weighted.lua, train_deepernet.lua, test*.lua, demo.lua of MKimiSH/video-filler).

NOT EXECUTED IN THIS REPOSITORY'S CI: the build image has no LuaJIT / Torch7 (SURVEY.md 8c).  It is the binding a
maintainer adds on a machine that has them; every C entry point it calls is declared in include/cenn.h and exercised by
the Python ctypes mirror (video-filler_b200/nn.py) that the parity tests drive.  INTEGRATION.md walks through it.

What it does
  * `require 'cunn'` / `require 'cutorch'` stand-ins: cutorch.setDevice(i) -> cenn_init(i-1); tensor:cuda() gives a
    torch.CudaTensor-like userdata (device pointer + sizes) backed by cenn_malloc / cenn_copy_h2d.
  * re-points the THNN entry of every module class on the hot path: nn.SpatialConvolution, nn.SpatialFullConvolution,
    nn.SpatialBatchNormalization, nn.LeakyReLU, nn.ReLU (nn.Threshold), nn.Tanh, nn.Sigmoid, nn.BCECriterion,
    nn.MSECriterion, nn.AbsCriterion, plus fused nn.MaskedMSECriterion and nn.GDLCriterion.  Class type names are
    kept ('nn.SpatialConvolution', ...) so weights_init (train.lua:58-67), the per-closure bias zeroing
    (train.lua:279-280) and util.save (util.lua:33-50) keep working.
  * optional fast path: cenn.Trainer wraps the whole-step executor (cenn_trainer_*) behind the two closures.
]]
local ffi = require 'ffi'

ffi.cdef[[
typedef struct cenn_state cenn_state;
typedef struct cenn_trainer cenn_trainer;
int cenn_init(int device, cenn_state **out);
int cenn_shutdown(cenn_state *s);
const char *cenn_last_error(void);
int cenn_set_precision(cenn_state *s, int mode);
int cenn_synchronize(cenn_state *s);
int cenn_malloc(cenn_state *s, size_t bytes, void **dptr);
int cenn_free(cenn_state *s, void *dptr);
int cenn_copy_h2d(cenn_state *s, void *dst, const void *src_host, size_t bytes);
int cenn_copy_d2h(cenn_state *s, void *dst_host, const void *src, size_t bytes);
int cenn_copy_d2d(cenn_state *s, void *dst, const void *src, size_t bytes);
int cenn_fill(cenn_state *s, float *x, int64_t n, float v);
int cenn_mul(cenn_state *s, float *x, int64_t n, float a);
int cenn_axpy(cenn_state *s, float *y, const float *x, int64_t n, float a);
int cenn_addcmul(cenn_state *s, float *y, float a, const float *p, const float *q, int64_t n);
int cenn_addcdiv(cenn_state *s, float *y, float a, const float *p, const float *q, int64_t n);
int cenn_sqrt(cenn_state *s, float *x, int64_t n);
int cenn_normal(cenn_state *s, float *x, int64_t n, float mean, float std, uint64_t seed);
int cenn_uniform(cenn_state *s, float *x, int64_t n, float a, float b, uint64_t seed);
int cenn_add_scalar(cenn_state *s, float *x, int64_t n, float a);
int cenn_cmul(cenn_state *s, float *y, const float *x, int64_t n);
int cenn_u8_to_float(cenn_state *s, float *dst, const uint8_t *src, int64_t n);
int cenn_masked_fill(cenn_state *s, float *x, const float *mask, int64_t n, float v);
int cenn_fill_box(cenn_state *s, float *x, int64_t N, int64_t C, int64_t H, int64_t W, int64_t c0, int64_t c1, int64_t y0, int64_t y1, int64_t x0, int64_t x1, float v);
int cenn_crop(cenn_state *s, float *dst, const float *src, int64_t N, int64_t C, int64_t H, int64_t W, int64_t y0, int64_t x0, int64_t h, int64_t w);
int cenn_MaskComposite(cenn_state *s, float *dst, const float *mask, const float *src, int64_t n);
int cenn_SpatialConvolutionMM_updateOutput(cenn_state *s, const float *input, float *output, const float *weight, const float *bias,
    int64_t batch, int64_t nInputPlane, int64_t inH, int64_t inW, int64_t nOutputPlane, int kW, int kH, int dW, int dH, int padW, int padH);
int cenn_SpatialConvolutionMM_updateGradInput(cenn_state *s, const float *gradOutput, float *gradInput, const float *weight,
    int64_t batch, int64_t nInputPlane, int64_t inH, int64_t inW, int64_t nOutputPlane, int kW, int kH, int dW, int dH, int padW, int padH);
int cenn_SpatialConvolutionMM_accGradParameters(cenn_state *s, const float *input, const float *gradOutput, float *gradWeight, float *gradBias,
    int64_t batch, int64_t nInputPlane, int64_t inH, int64_t inW, int64_t nOutputPlane, int kW, int kH, int dW, int dH, int padW, int padH, float scale);
int cenn_SpatialFullConvolution_updateOutput(cenn_state *s, const float *input, float *output, const float *weight, const float *bias,
    int64_t batch, int64_t nInputPlane, int64_t inH, int64_t inW, int64_t nOutputPlane, int kW, int kH, int dW, int dH, int padW, int padH, int adjW, int adjH);
int cenn_SpatialFullConvolution_updateGradInput(cenn_state *s, const float *gradOutput, float *gradInput, const float *weight,
    int64_t batch, int64_t nInputPlane, int64_t inH, int64_t inW, int64_t nOutputPlane, int kW, int kH, int dW, int dH, int padW, int padH, int adjW, int adjH);
int cenn_SpatialFullConvolution_accGradParameters(cenn_state *s, const float *input, const float *gradOutput, float *gradWeight, float *gradBias,
    int64_t batch, int64_t nInputPlane, int64_t inH, int64_t inW, int64_t nOutputPlane, int kW, int kH, int dW, int dH, int padW, int padH, int adjW, int adjH, float scale);
int cenn_BatchNormalization_updateOutput(cenn_state *s, const float *input, float *output, const float *weight, const float *bias,
    float *runningMean, float *runningVar, float *saveMean, float *saveStd, int64_t batch, int64_t C, int64_t spatial, int train, double momentum, double eps);
int cenn_BatchNormalization_backward(cenn_state *s, const float *input, const float *gradOutput, float *gradInput, float *gradWeight, float *gradBias,
    const float *weight, const float *runningMean, const float *runningVar, const float *saveMean, const float *saveStd,
    int64_t batch, int64_t C, int64_t spatial, int train, double scale, double eps);
int cenn_LeakyReLU_updateOutput(cenn_state *s, const float *input, float *output, int64_t n, double negval, int inplace);
int cenn_LeakyReLU_updateGradInput(cenn_state *s, const float *input, const float *gradOutput, float *gradInput, int64_t n, double negval, int inplace);
int cenn_Threshold_updateOutput(cenn_state *s, const float *input, float *output, int64_t n, double threshold, double val, int inplace);
int cenn_Threshold_updateGradInput(cenn_state *s, const float *input, const float *gradOutput, float *gradInput, int64_t n, double threshold, int inplace);
int cenn_Tanh_updateOutput(cenn_state *s, const float *input, float *output, int64_t n);
int cenn_Tanh_updateGradInput(cenn_state *s, const float *gradOutput, float *gradInput, const float *output, int64_t n);
int cenn_Sigmoid_updateOutput(cenn_state *s, const float *input, float *output, int64_t n);
int cenn_Sigmoid_updateGradInput(cenn_state *s, const float *gradOutput, float *gradInput, const float *output, int64_t n);
int cenn_BCECriterion_updateOutput(cenn_state *s, const float *input, const float *target, int64_t n, int sizeAverage, float *loss_host);
int cenn_BCECriterion_updateGradInput(cenn_state *s, const float *input, const float *target, float *gradInput, int64_t n, int sizeAverage);
int cenn_MSECriterion_updateOutput(cenn_state *s, const float *input, const float *target, int64_t n, int sizeAverage, float *loss_host);
int cenn_MSECriterion_updateGradInput(cenn_state *s, const float *input, const float *target, float *gradInput, int64_t n, int sizeAverage);
int cenn_MaskedMSECriterion_forward_backward(cenn_state *s, const float *input, const float *target, const float *mask, float *gradInput,
    int64_t n, double mWeight, float *loss_host);
int cenn_GDLCriterion_forward_backward(cenn_state *s, const float *input, const float *target, float *gradInput,
    int64_t batch, int64_t C, int64_t H, int64_t W, float *loss_host);
int cenn_AdamFlat(cenn_state *s, float *x, const float *g, float *m, float *v, int64_t n, double lr, double beta1, double beta2, double eps, int64_t t);
int cenn_JoinTable_updateOutput(cenn_state *s, float *joined, const float *part, int64_t batch, int64_t joined_per_sample, int64_t offset, int64_t part_per_sample);
int cenn_JoinTable_updateGradInput(cenn_state *s, const float *gradJoined, float *gradPart, int64_t batch, int64_t joined_per_sample, int64_t offset, int64_t part_per_sample);
]]

local lib = ffi.load(os.getenv('CENN_LIB') or 'libcenn.so')
local cenn = { lib = lib, state = nil }

-- THNN functions return void and raise through THError; libcenn returns a status and a thread-local message.
local function check(rc) if rc ~= 0 then error(ffi.string(lib.cenn_last_error()), 3) end end

---------------------------------------------------------------------------------------------- cutorch stand-in
local cutorch = {}
function cutorch.setDevice(i)                       -- train.lua:250 (1-based)
  local out = ffi.new('cenn_state*[1]')
  check(lib.cenn_init(i - 1, out)); cenn.state = out[0]
end
function cutorch.synchronize() check(lib.cenn_synchronize(cenn.state)) end
package.loaded['cutorch'] = cutorch
_G.cutorch = cutorch

---------------------------------------------------------------------------------------------- CudaTensor (fp32, contiguous)
local CudaTensor = {}; CudaTensor.__index = CudaTensor
local function numel(sz) local n = 1; for _, d in ipairs(sz) do n = n * d end; return n end
function cenn.CudaTensor(...)
  local sz = {...}; if type(sz[1]) == 'table' then sz = sz[1] end
  local p = ffi.new('void*[1]')
  check(lib.cenn_malloc(cenn.state, math.max(1, numel(sz)) * 4, p))
  local t = setmetatable({ sz = sz, ptr = ffi.cast('float*', p[0]) }, CudaTensor)
  t.gc = ffi.gc(p[0], function(q) lib.cenn_free(cenn.state, q) end)
  return t
end
function CudaTensor:nElement() return numel(self.sz) end
function CudaTensor:size(d) if d then return self.sz[d] end; return torch.LongStorage(self.sz) end
function CudaTensor:dim() return #self.sz end
function CudaTensor:nDimension() return #self.sz end
function CudaTensor:type(ty)                            -- tensor:type() / tensor:type('torch.FloatTensor') / :typeAs
  if not ty then return 'torch.CudaTensor' end
  if ty == 'torch.CudaTensor' then return self end
  local t = self:float(); return ty == 'torch.FloatTensor' and t or t:type(ty)
end
function CudaTensor:typeAs(o) return self:type(torch.type(o) == 'table' and 'torch.CudaTensor' or torch.type(o)) end
function CudaTensor:cuda() return self end
function CudaTensor:contiguous() return self end          -- stand-in tensors are always dense
function CudaTensor:isContiguous() return true end
function CudaTensor:isSameSizeAs(o) if #self.sz ~= #o.sz then return false end; for i, d in ipairs(self.sz) do if d ~= o.sz[i] then return false end end; return true end
function CudaTensor:fill(v) check(lib.cenn_fill(cenn.state, self.ptr, self:nElement(), v)); return self end
function CudaTensor:zero() return self:fill(0) end
function CudaTensor:mul(a) check(lib.cenn_mul(cenn.state, self.ptr, self:nElement(), a)); return self end
function CudaTensor:div(a) return self:mul(1 / a) end
-- x:add(s) | x:add(t) | x:add(s, t)   (train.lua:398; train_vid_weighted.lua:526; optim.adam)
function CudaTensor:add(a, x)
  if x == nil and type(a) == 'number' then check(lib.cenn_add_scalar(cenn.state, self.ptr, self:nElement(), a)); return self end
  if x == nil then a, x = 1, a end
  check(lib.cenn_axpy(cenn.state, self.ptr, x.ptr, self:nElement(), a)); return self
end
function CudaTensor:cmul(x) check(lib.cenn_cmul(cenn.state, self.ptr, x.ptr, self:nElement())); return self end              -- train_vid_weighted.lua:496
-- y:addcmul([a,] p, q) / y:addcdiv([a,] p, q)   (train.lua:394; optim.adam: v:addcmul(1-beta2, g, g), x:addcdiv(-step, m, denom))
function CudaTensor:addcmul(a, p, q) if q == nil then a, p, q = 1, a, p end; check(lib.cenn_addcmul(cenn.state, self.ptr, a, p.ptr, q.ptr, self:nElement())); return self end
function CudaTensor:addcdiv(a, p, q) if q == nil then a, p, q = 1, a, p end; check(lib.cenn_addcdiv(cenn.state, self.ptr, a, p.ptr, q.ptr, self:nElement())); return self end
function CudaTensor:sqrt() check(lib.cenn_sqrt(cenn.state, self.ptr, self:nElement())); return self end
local seed_ctr = 0
local function next_seed() seed_ctr = seed_ctr + 1; return (torch.initialSeed() % 2147483647) * 4096 + seed_ctr end          -- one Philox stream per call
function CudaTensor:normal(mean, std) check(lib.cenn_normal(cenn.state, self.ptr, self:nElement(), mean or 0, std or 1, next_seed())); return self end   -- train.lua:61,64,272
function CudaTensor:uniform(a, b) check(lib.cenn_uniform(cenn.state, self.ptr, self:nElement(), a or 0, b or 1, next_seed())); return self end        -- train.lua:270
function CudaTensor:copy(src)                         -- from a host tensor (H2D; ByteTensor masks become {0,1} floats) or a CudaTensor / window (D2D)
  if getmetatable(src) == CudaTensor then check(lib.cenn_copy_d2d(cenn.state, self.ptr, src.ptr, self:nElement() * 4))
  elseif torch.type(src) == 'torch.ByteTensor' then            -- input_mask:copy(real_mask) (train_vid_weighted.lua:391)
    src = src:contiguous()
    local p = ffi.new('void*[1]'); check(lib.cenn_malloc(cenn.state, self:nElement(), p))
    check(lib.cenn_copy_h2d(cenn.state, p[0], src:data(), self:nElement()))
    check(lib.cenn_u8_to_float(cenn.state, self.ptr, ffi.cast('const uint8_t*', p[0]), self:nElement())); check(lib.cenn_free(cenn.state, p[0]))
  else src = src:float():contiguous(); check(lib.cenn_copy_h2d(cenn.state, self.ptr, src:data(), self:nElement() * 4)) end
  return self
end
function CudaTensor:float()
  local t = torch.FloatTensor(unpack(self.sz))
  check(lib.cenn_copy_d2h(cenn.state, t:data(), self.ptr, self:nElement() * 4)); return t
end
function CudaTensor:double() return self:float():double() end
function CudaTensor:byte() return self:float():byte() end
function CudaTensor:clone() return cenn.CudaTensor(self.sz):copy(self) end                                   -- train.lua:287,389
function CudaTensor.new(self_or_size, ...)                -- x.new(size) / x.new() (optim.adam state, nn module buffers)
  local a = {...}
  if getmetatable(self_or_size) == CudaTensor then
    if #a == 0 then return cenn.CudaTensor(0) end
    if type(a[1]) == 'number' then return cenn.CudaTensor(a) end
    return cenn.CudaTensor(a[1].totable and a[1]:totable() or a[1])
  end
  return cenn.CudaTensor(self_or_size, ...)
end
function CudaTensor:resize(...) local sz = {...}
  if type(sz[1]) ~= 'number' then sz = sz[1].totable and sz[1]:totable() or sz[1] end
  if numel(sz) ~= self:nElement() then local n = cenn.CudaTensor(sz); self.ptr, self.gc = n.ptr, n.gc end
  self.sz = sz; return self
end
function CudaTensor:resizeAs(o) return self:resize(unpack(o.sz or o:size():totable())) end
function CudaTensor:view(...)                           -- shares storage (train_vid_weighted.lua:163 mask:view; nn.View)
  local sz = {...}; if type(sz[1]) ~= 'number' then sz = sz[1]:totable() end
  local known, neg = 1, nil
  for i, d in ipairs(sz) do if d == -1 then neg = i else known = known * d end end
  if neg then sz[neg] = self:nElement() / known end
  assert(numel(sz) == self:nElement(), 'view: size mismatch')
  return setmetatable({ sz = sz, ptr = self.ptr, gc = self.gc }, CudaTensor)
end
function CudaTensor:viewAs(o) return self:view(unpack(o.sz)) end
function CudaTensor:narrow(dim, first, n)               -- outermost dimension only (getParameters views, batch slices): shares storage
  assert(dim == 1, 'narrow: only the first dimension of a dense tensor can be narrowed without a copy')
  local sz = {unpack(self.sz)}; sz[1] = n
  return setmetatable({ sz = sz, ptr = self.ptr + (first - 1) * (self:nElement() / self.sz[1]), gc = self.gc }, CudaTensor)
end
-- mask arithmetic of the video scripts.  Masks are {0,1} float tensors once on the device (train_vid_weighted.lua:302,391).
function CudaTensor:maskedFill(mask, v) check(lib.cenn_masked_fill(cenn.state, self.ptr, mask.ptr, self:nElement(), v)); return self end    -- inpaint_utils.lua:45-58
-- x:maskedSelect(mask) is only ever consumed by y:maskedCopy(mask, sel) with the SAME mask (train_vid_weighted.lua:430-434,
-- inpaint_utils.lua:79-90): the pair is a composite y = where(mask, x, y), so the selection stays lazy and no compaction runs.
function CudaTensor:maskedSelect(mask) return { lazy_select = true, src = self, mask = mask } end
function CudaTensor:maskedCopy(mask, sel)
  assert(type(sel) == 'table' and sel.lazy_select and sel.mask == mask, 'maskedCopy: expects the result of maskedSelect with the same mask')
  check(lib.cenn_MaskComposite(cenn.state, self.ptr, mask.ptr, sel.src.ptr, self:nElement())); return self
end
-- reductions are only used for logging / display in the scripts: computed on the host copy
function CudaTensor:min() return self:float():min() end
function CudaTensor:max() return self:float():max() end
function CudaTensor:mean() return self:float():mean() end
function CudaTensor:std() return self:float():std() end
function CudaTensor:sum() return self:float():sum() end
function CudaTensor:norm(p) return self:float():norm(p) end
-- range indexing t[{{},{},{a,b},{c,d}}] (train.lua:287-290,392): a window object that supports :fill / :copy / :clone
local Window = {}; Window.__index = Window
local function bounds(t, idx)
  assert(#t.sz == 4, 'range indexing is implemented for [N,C,H,W] tensors')
  local b = {}
  for d = 1, 4 do local r = idx[d] or {}; if type(r) == 'number' then r = {r, r} end; b[d] = { (r[1] or 1) - 1, r[2] or t.sz[d] } end
  assert(b[1][1] == 0 and b[1][2] == t.sz[1], 'range indexing keeps the whole batch dimension')
  return b
end
function Window:fill(v) local t, b = self.t, self.b
  check(lib.cenn_fill_box(cenn.state, t.ptr, t.sz[1], t.sz[2], t.sz[3], t.sz[4], b[2][1], b[2][2], b[3][1], b[3][2], b[4][1], b[4][2], v)); return self end
function Window:clone() local t, b = self.t, self.b
  assert(b[2][1] == 0 and b[2][2] == t.sz[2], 'window clone keeps all channels')
  local out = cenn.CudaTensor(t.sz[1], t.sz[2], b[3][2] - b[3][1], b[4][2] - b[4][1])
  check(lib.cenn_crop(cenn.state, out.ptr, t.ptr, t.sz[1], t.sz[2], t.sz[3], t.sz[4], b[3][1], b[4][1], b[3][2] - b[3][1], b[4][2] - b[4][1])); return out end
function Window:copy(src) error('window:copy is not used by the scripts on GPU tensors') end
-- an index that restricts only the FIRST dimension (dst[{{b,b+step-1},{},{}}], dst[b]; inpaint_utils.lua:49-56,93-97) is a dense slice:
-- it becomes a narrow() view with every tensor method; a 4-D box becomes a Window
local function first_dim_only(t, k)
  for d = 2, #t.sz do local r = k[d]; if r ~= nil and (type(r) == 'number' or r[1] ~= nil) then return false end end
  return true
end
CudaTensor.__index = function(t, k)
  if type(k) == 'number' then local v = t:narrow(1, k, 1); local sz = {unpack(t.sz)}; table.remove(sz, 1); v.sz = sz; return v end   -- dst[b]
  if type(k) == 'table' then
    if first_dim_only(t, k) then local r = k[1] or {}; if type(r) == 'number' then r = {r, r} end
      local a, b = r[1] or 1, r[2] or t.sz[1]; return t:narrow(1, a, b - a + 1) end
    return setmetatable({ t = t, b = bounds(t, k) }, Window)
  end
  return CudaTensor[k]
end
CudaTensor.__newindex = function(t, k, v)                -- t[{...}] = scalar
  if type(k) == 'table' then setmetatable({ t = t, b = bounds(t, k) }, Window):fill(v) else rawset(t, k, v) end
end
torch.FloatTensor.cuda = function(self) return cenn.CudaTensor(self:size():totable()):copy(self) end
torch.ByteTensor.cuda = function(self) return cenn.CudaTensor(self:size():totable()):copy(self) end

require 'nn'
-- module:cuda() / module:float() (train.lua:254; util.lua:76,81): stock nn.Module:type converts through torch.Tensor.type, which the
-- stand-in tensor is not part of -- walk the module tree and convert every tensor field in place (shared tensors stay shared)
local function convert_tree(obj, to_cuda, seen)
  seen = seen or {}
  if type(obj) ~= 'table' or seen[obj] then return obj end
  seen[obj] = true
  for k, v in pairs(obj) do
    if to_cuda and torch.isTensor(v) and torch.type(v) ~= 'torch.CudaTensor' then
      seen[v] = seen[v] or (torch.type(v) == 'torch.LongTensor' and v or v:cuda()); obj[k] = seen[v]
    elseif not to_cuda and getmetatable(v) == CudaTensor then
      seen[v] = seen[v] or v:float(); obj[k] = seen[v]
    elseif type(v) == 'table' and getmetatable(v) ~= CudaTensor then convert_tree(v, to_cuda, seen) end
  end
  return obj
end
function nn.Module:cuda() return convert_tree(self, true) end
function nn.Module:float() return convert_tree(self, false) end
function nn.Criterion:cuda() return convert_tree(self, true) end
function nn.Criterion:float() return convert_tree(self, false) end

---------------------------------------------------------------------------------------------- nn modules: THNN -> libcenn
-- Each updateOutput / updateGradInput / accGradParameters below replaces the `input.THNN.<Op>_<phase>(...)` call of
-- the stock Lua method (same argument meaning; THCudaTensor* becomes pointer + explicit sizes).
require 'nn'
local S = function() return cenn.state end

local Conv = nn.SpatialConvolution
function Conv:updateOutput(input)                   -- THNN SpatialConvolutionMM_updateOutput (train.lua:89-104,183-196)
  local N, C, H, W = unpack(input.sz)
  local oH = math.floor((H + 2 * self.padH - self.kH) / self.dH) + 1
  local oW = math.floor((W + 2 * self.padW - self.kW) / self.dW) + 1
  self.output:resize(N, self.nOutputPlane, oH, oW)
  check(lib.cenn_SpatialConvolutionMM_updateOutput(S(), input.ptr, self.output.ptr, self.weight.ptr, self.bias and self.bias.ptr,
        N, C, H, W, self.nOutputPlane, self.kW, self.kH, self.dW, self.dH, self.padW, self.padH))
  return self.output
end
function Conv:updateGradInput(input, gradOutput)
  local N, C, H, W = unpack(input.sz)
  self.gradInput:resizeAs(input)
  check(lib.cenn_SpatialConvolutionMM_updateGradInput(S(), gradOutput.ptr, self.gradInput.ptr, self.weight.ptr,
        N, C, H, W, self.nOutputPlane, self.kW, self.kH, self.dW, self.dH, self.padW, self.padH))
  return self.gradInput
end
function Conv:accGradParameters(input, gradOutput, scale)
  local N, C, H, W = unpack(input.sz)
  check(lib.cenn_SpatialConvolutionMM_accGradParameters(S(), input.ptr, gradOutput.ptr, self.gradWeight.ptr, self.gradBias and self.gradBias.ptr,
        N, C, H, W, self.nOutputPlane, self.kW, self.kH, self.dW, self.dH, self.padW, self.padH, scale or 1))
end

local Full = nn.SpatialFullConvolution               -- train.lua:134-146; weight [nInputPlane][nOutputPlane][kH][kW]
function Full:updateOutput(input)
  local N, C, H, W = unpack(input.sz)
  local oH = (H - 1) * self.dH - 2 * self.padH + self.kH + self.adjH
  local oW = (W - 1) * self.dW - 2 * self.padW + self.kW + self.adjW
  self.output:resize(N, self.nOutputPlane, oH, oW)
  check(lib.cenn_SpatialFullConvolution_updateOutput(S(), input.ptr, self.output.ptr, self.weight.ptr, self.bias and self.bias.ptr,
        N, C, H, W, self.nOutputPlane, self.kW, self.kH, self.dW, self.dH, self.padW, self.padH, self.adjW, self.adjH))
  return self.output
end
function Full:updateGradInput(input, gradOutput)
  local N, C, H, W = unpack(input.sz)
  self.gradInput:resizeAs(input)
  check(lib.cenn_SpatialFullConvolution_updateGradInput(S(), gradOutput.ptr, self.gradInput.ptr, self.weight.ptr,
        N, C, H, W, self.nOutputPlane, self.kW, self.kH, self.dW, self.dH, self.padW, self.padH, self.adjW, self.adjH))
  return self.gradInput
end
function Full:accGradParameters(input, gradOutput, scale)
  local N, C, H, W = unpack(input.sz)
  check(lib.cenn_SpatialFullConvolution_accGradParameters(S(), input.ptr, gradOutput.ptr, self.gradWeight.ptr, self.gradBias and self.gradBias.ptr,
        N, C, H, W, self.nOutputPlane, self.kW, self.kH, self.dW, self.dH, self.padW, self.padH, self.adjW, self.adjH, scale or 1))
end

local BN = nn.BatchNormalization                     -- nn.SpatialBatchNormalization inherits (train.lua:79)
function BN:updateOutput(input)
  local N, C = input.sz[1], input.sz[2]
  local spatial = input:nElement() / (N * C)
  self.output:resizeAs(input); self.save_mean = self.save_mean or cenn.CudaTensor(C); self.save_std = self.save_std or cenn.CudaTensor(C)
  check(lib.cenn_BatchNormalization_updateOutput(S(), input.ptr, self.output.ptr, self.weight and self.weight.ptr, self.bias and self.bias.ptr,
        self.running_mean.ptr, self.running_var.ptr, self.save_mean.ptr, self.save_std.ptr, N, C, spatial, self.train and 1 or 0, self.momentum, self.eps))
  return self.output
end
local function bn_backward(self, input, gradOutput, scale, gradInput, gradWeight, gradBias)
  local N, C = input.sz[1], input.sz[2]
  check(lib.cenn_BatchNormalization_backward(S(), input.ptr, gradOutput.ptr, gradInput and gradInput.ptr, gradWeight and gradWeight.ptr,
        gradBias and gradBias.ptr, self.weight and self.weight.ptr, self.running_mean.ptr, self.running_var.ptr, self.save_mean.ptr,
        self.save_std.ptr, N, C, input:nElement() / (N * C), self.train and 1 or 0, scale or 1, self.eps))
end
function BN:backward(input, gradOutput, scale) self.gradInput:resizeAs(input); bn_backward(self, input, gradOutput, scale, self.gradInput, self.gradWeight, self.gradBias); return self.gradInput end
function BN:updateGradInput(input, gradOutput) self.gradInput:resizeAs(input); bn_backward(self, input, gradOutput, 1, self.gradInput); return self.gradInput end
function BN:accGradParameters(input, gradOutput, scale) bn_backward(self, input, gradOutput, scale, nil, self.gradWeight, self.gradBias) end

function nn.LeakyReLU:updateOutput(input)
  if not self.inplace then self.output:resizeAs(input) else self.output = input end
  check(lib.cenn_LeakyReLU_updateOutput(S(), input.ptr, self.output.ptr, input:nElement(), self.negval, self.inplace and 1 or 0)); return self.output
end
function nn.LeakyReLU:updateGradInput(input, gradOutput)
  if not self.inplace then self.gradInput:resizeAs(input) else self.gradInput = gradOutput end
  check(lib.cenn_LeakyReLU_updateGradInput(S(), input.ptr, gradOutput.ptr, self.gradInput.ptr, input:nElement(), self.negval, self.inplace and 1 or 0)); return self.gradInput
end
function nn.Threshold:updateOutput(input)            -- nn.ReLU = nn.Threshold(0, 0, inplace)
  if not self.inplace then self.output:resizeAs(input) else self.output = input end
  check(lib.cenn_Threshold_updateOutput(S(), input.ptr, self.output.ptr, input:nElement(), self.threshold, self.val, self.inplace and 1 or 0)); return self.output
end
function nn.Threshold:updateGradInput(input, gradOutput)
  if not self.inplace then self.gradInput:resizeAs(input) else self.gradInput = gradOutput end
  check(lib.cenn_Threshold_updateGradInput(S(), input.ptr, gradOutput.ptr, self.gradInput.ptr, input:nElement(), self.threshold, self.inplace and 1 or 0)); return self.gradInput
end
function nn.Tanh:updateOutput(input) self.output:resizeAs(input); check(lib.cenn_Tanh_updateOutput(S(), input.ptr, self.output.ptr, input:nElement())); return self.output end
function nn.Tanh:updateGradInput(input, gradOutput) self.gradInput:resizeAs(input); check(lib.cenn_Tanh_updateGradInput(S(), gradOutput.ptr, self.gradInput.ptr, self.output.ptr, input:nElement())); return self.gradInput end
function nn.Sigmoid:updateOutput(input) self.output:resizeAs(input); check(lib.cenn_Sigmoid_updateOutput(S(), input.ptr, self.output.ptr, input:nElement())); return self.output end
function nn.Sigmoid:updateGradInput(input, gradOutput) self.gradInput:resizeAs(input); check(lib.cenn_Sigmoid_updateGradInput(S(), gradOutput.ptr, self.gradInput.ptr, self.output.ptr, input:nElement())); return self.gradInput end

-- nn.JoinTable(2) of the noiseGen / conditionAdv nets (train.lua:120,177): stock nn.JoinTable narrows and copies through tensor
-- methods the stand-in tensor does not have; here each table member is one strided device copy.  nn.ParallelTable is pure Lua.
local function per_sample(t) local n = 1; for d = 2, t:dim() do n = n * t.sz[d] end; return n end
function nn.JoinTable:updateOutput(input)
  assert(self.dimension == 2, 'JoinTable: the scripts join along dimension 2 (channels) only')
  local sz = {unpack(input[1].sz)}; sz[2] = 0
  for _, x in ipairs(input) do sz[2] = sz[2] + x.sz[2] end
  self.output = self.output.ptr and self.output or cenn.CudaTensor(1); self.output:resize(unpack(sz))
  local total, off = per_sample(self.output), 0
  for _, x in ipairs(input) do
    check(lib.cenn_JoinTable_updateOutput(S(), self.output.ptr, x.ptr, sz[1], total, off, per_sample(x))); off = off + per_sample(x)
  end
  return self.output
end
function nn.JoinTable:updateGradInput(input, gradOutput)
  local total, off = per_sample(gradOutput), 0
  self.gradInput = type(self.gradInput) == 'table' and self.gradInput or {}
  for i, x in ipairs(input) do
    self.gradInput[i] = self.gradInput[i] or cenn.CudaTensor(1); self.gradInput[i]:resizeAs(x)
    check(lib.cenn_JoinTable_updateGradInput(S(), gradOutput.ptr, self.gradInput[i].ptr, x.sz[1], total, off, per_sample(x))); off = off + per_sample(x)
  end
  return self.gradInput
end

---------------------------------------------------------------------------------------------- criteria
local loss = ffi.new('float[1]')
function nn.BCECriterion:updateOutput(input, target)  -- returns a Lua number (the only mandatory sync, SURVEY.md 8b)
  check(lib.cenn_BCECriterion_updateOutput(S(), input.ptr, target.ptr, input:nElement(), self.sizeAverage and 1 or 0, loss)); self.output = loss[0]; return self.output
end
function nn.BCECriterion:updateGradInput(input, target)
  self.gradInput:resizeAs(input); check(lib.cenn_BCECriterion_updateGradInput(S(), input.ptr, target.ptr, self.gradInput.ptr, input:nElement(), self.sizeAverage and 1 or 0)); return self.gradInput
end
function nn.MSECriterion:updateOutput(input, target)
  check(lib.cenn_MSECriterion_updateOutput(S(), input.ptr, target.ptr, input:nElement(), self.sizeAverage and 1 or 0, loss)); self.output = loss[0]; return self.output
end
function nn.MSECriterion:updateGradInput(input, target)
  self.gradInput:resizeAs(input); check(lib.cenn_MSECriterion_updateGradInput(S(), input.ptr, target.ptr, self.gradInput.ptr, input:nElement(), self.sizeAverage and 1 or 0)); return self.gradInput
end

-- repo-local criteria (MaskedMSECriterion.lua:4-42, gdl_criterion.lua:4-53): one fused kernel each instead of an nngraph
local MaskedMSE, mparent = torch.class('nn.MaskedMSECriterion', 'nn.Criterion')
function MaskedMSE:__init(mWeight) mparent.__init(self); assert(mWeight, 'mWeight required (MaskedMSECriterion.lua:15)'); self.mWeight = mWeight; self.gradInput = cenn.CudaTensor(1) end
function MaskedMSE:setMask(m) assert(torch.type(m) == 'torch.ByteTensor'); self.mask = m:float():cuda() end   -- MaskedMSECriterion.lua:24-27
function MaskedMSE:updateOutput(input, target)
  self.gradInput:resizeAs(input)
  check(lib.cenn_MaskedMSECriterion_forward_backward(S(), input.ptr, target.ptr, self.mask.ptr, self.gradInput.ptr, input:nElement(), self.mWeight, loss)); self.output = loss[0]; return self.output
end
function MaskedMSE:updateGradInput() return self.gradInput end
local GDL, gparent = torch.class('nn.GDLCriterion', 'nn.Criterion')
function GDL:__init(alpha) gparent.__init(self); assert((alpha or 1) == 1, 'alpha must be 1 (gdl_criterion.lua:9)'); self.gradInput = cenn.CudaTensor(1) end
function GDL:updateOutput(input, target)
  local N, C, H, W = unpack(input.sz); self.gradInput:resizeAs(input)
  check(lib.cenn_GDLCriterion_forward_backward(S(), input.ptr, target.ptr, self.gradInput.ptr, N, C, H, W, loss)); self.output = loss[0]; return self.output
end
function GDL:updateGradInput() return self.gradInput end

---------------------------------------------------------------------------------------------- optional fast paths
-- Both are opt-in: the unchanged scripts never touch them.  A maintainer selects them with two lines in train.lua
-- (`local fast = cenn.Trainer(opt, netG, netD)` ... `fast:step(real_ctx, real_center)`) or in test_vid_wholeim.lua
-- (`local eng = cenn.Inpainter(opt, net)` ... `eng:sweep(images01, mask)`).
ffi.cdef[[
typedef struct cenn_trainer_config { int variant, batchSize, fineSize, nBottleneck, nef, ngf, ndf, nc, predLen, overlapPred;
    float wtl2, weight_nomask, wtgdl, lr, beta1; int precision, world_size, rank, dead_dgrad, noiseGen, nz, conditionAdv, bn_local; } cenn_trainer_config;
int cenn_trainer_create(cenn_state *s, const cenn_trainer_config *cfg, cenn_trainer **out);
int cenn_trainer_destroy(cenn_trainer *t);
int cenn_trainer_param_count(cenn_trainer *t, int net, int64_t *count);
int cenn_trainer_set_params_host(cenn_trainer *t, int net, const float *flat_host);
int cenn_trainer_get_params_host(cenn_trainer *t, int net, float *flat_host);
int cenn_trainer_bn_stat_count(cenn_trainer *t, int net, int64_t *count);
int cenn_trainer_set_bn_stats_host(cenn_trainer *t, int net, const float *stats_host);
int cenn_trainer_get_bn_stats_host(cenn_trainer *t, int net, float *stats_host);
int cenn_trainer_step_host(cenn_trainer *t, const float *a_host, const float *b_host, const uint8_t *mask_host, float *losses_host);
int cenn_trainer_step_host_async(cenn_trainer *t, const float *a_host, const float *b_host, const uint8_t *mask_host);
int cenn_trainer_wait_losses(cenn_trainer *t, float *losses_host);
int cenn_trainer_step_clips_host(cenn_trainer *t, const float *frames01_host, const uint8_t *mask1_host, const uint8_t *flip_host, float maskValue, float *losses_host);
int cenn_trainer_step_frames_host(cenn_trainer *t, const uint8_t *frames_u8_host, int iH, int iW, const uint8_t *mask_full_host,
    const int *crop_host, const uint8_t *flip_host, const int *blocks_host, float maskValue, float *losses_host);
int cenn_trainer_step_images_u8_host(cenn_trainer *t, const uint8_t *images_u8_host, float *losses_host);
int cenn_trainer_set_noise_host(cenn_trainer *t, const float *noise_host);
typedef struct cenn_inpainter cenn_inpainter;
typedef struct cenn_inpainter_config { int variant, batch, fineSize, nBottleneck, nef, ngf, nc, inputLen; } cenn_inpainter_config;
int cenn_inpainter_create(cenn_state *s, const cenn_inpainter_config *cfg, cenn_inpainter **out);
int cenn_inpainter_destroy(cenn_inpainter *p);
int cenn_inpainter_param_count(cenn_inpainter *p, int64_t *params, int64_t *bn_stats);
int cenn_inpainter_load_host(cenn_inpainter *p, const float *flat_host, const float *bn_stats_host);
int cenn_inpainter_forward_host(cenn_inpainter *p, const float *in_host, float *out_host, int n);
int cenn_inpainter_sweep_host(cenn_inpainter *p, cenn_inpainter *init, const float *frames01_host, const uint8_t *mask_host,
    int P, int inh, int inw, float maskValue, float *out01_host, float *full01_host, float *inpaint01_host);
]]

-- [running_mean, running_var] of every BN module in execution order, as one FloatTensor (what the *_bn_stats_* calls carry)
local function bn_stats_of(net)
  local parts = {}
  for _, m in ipairs(net:findModules('nn.SpatialBatchNormalization')) do
    parts[#parts + 1] = m.running_mean:float(); parts[#parts + 1] = (m.running_var or torch.pow(m.running_std, -2):add(-m.eps)):float()  -- util.lua:40-44
  end
  return torch.cat(parts)
end

-- whole-step executor behind fDx/fGx + the two optim.adam calls (train.lua:278-424; train_vid_weighted.lua:373-537)
local Trainer = torch.class('cenn.Trainer')
function Trainer:__init(opt, netG, netD, world_size, rank)
  local video = opt.predLen ~= nil
  local cfg = ffi.new('cenn_trainer_config', {video and 1 or 0, opt.batchSize, opt.fineSize, opt.nBottleneck, opt.nef, opt.ngf, opt.ndf, opt.nc or 3,
    opt.predLen or 1, opt.overlapPred or 0, opt.wtl2, opt.weight_nomask or 0, opt.wtgdl or 0, opt.lr, opt.beta1, 1, world_size or 1, rank or 0, 1,
    opt.noiseGen and 1 or 0, opt.nz or 100, opt.conditionAdv and 1 or 0, opt.bn_local and 1 or 0})
  local h = ffi.new('cenn_trainer*[1]'); check(lib.cenn_trainer_create(S(), cfg, h)); self.h = ffi.gc(h[0], lib.cenn_trainer_destroy)
  self.netG, self.netD = netG, netD
  self.pG, self.pD = netG:getParameters(), netD:getParameters()          -- train.lua:262-263: the flat vectors are the interchange format
  check(lib.cenn_trainer_set_params_host(self.h, 0, self.pG:float():data())); check(lib.cenn_trainer_set_params_host(self.h, 1, self.pD:float():data()))
  self.losses = torch.FloatTensor(8)
end
-- one G+D step; returns errD, errG, errG_l2 as printed at train.lua:443-450 (host FloatTensors in, like the data loader delivers them)
-- opt.noiseGen: pass the step's noise draw (train.lua:319-323) as the 4th argument, a FloatTensor [B, nz, 1, 1]
function Trainer:step(a, b, mask, noise)
  if noise then check(lib.cenn_trainer_set_noise_host(self.h, noise:data())) end
  check(lib.cenn_trainer_step_host(self.h, a:data(), b:data(), mask and mask:data() or nil, self.losses:data()))
  return self.losses[1], self.losses[2], self.losses[3]
end
-- byte inputs: the image loader's crop as a ByteTensor [B,3,F,F] (rescale, centre clone, mean fill of train.lua:286-290 on the device) ...
function Trainer:stepBytes(images_u8)
  check(lib.cenn_trainer_step_images_u8_host(self.h, images_u8:data(), self.losses:data())); return self.losses[1], self.losses[2], self.losses[3]
end
-- ... the video loader's sample after its crop (frames in [0,1], one mask plane per clip, hflip flags: datavid/donkey_folder.lua:161-187) ...
function Trainer:stepClips(frames01, mask1, flip, maskValue)
  check(lib.cenn_trainer_step_clips_host(self.h, frames01:data(), mask1:data(), flip and flip:data() or nil, maskValue, self.losses:data()))
  return self.losses[1], self.losses[2], self.losses[3]
end
-- ... or whole decoded frames plus the hook's random draws (crop origins IntTensor [B,2], flip ByteTensor [B], block tables IntTensor [B,21]):
-- crop, mask crop / random-block mask, maskedFill, hflip and rescale of datavid/donkey_folder.lua:114-129,138-187 on the device
function Trainer:stepFrames(frames_u8, mask_full, crop, flip, blocks, maskValue)
  check(lib.cenn_trainer_step_frames_host(self.h, frames_u8:data(), frames_u8:size(3), frames_u8:size(4), mask_full:data(), crop:data(),
        flip and flip:data() or nil, blocks:data(), maskValue, self.losses:data()))
  return self.losses[1], self.losses[2], self.losses[3]
end
-- write parameters and running statistics back into the nn modules (before util.save, util.lua:72-97)
function Trainer:sync()
  local f = torch.FloatTensor(self.pG:nElement()); check(lib.cenn_trainer_get_params_host(self.h, 0, f:data())); self.pG:copy(f)
  f = torch.FloatTensor(self.pD:nElement()); check(lib.cenn_trainer_get_params_host(self.h, 1, f:data())); self.pD:copy(f)
  for idx, net in ipairs{self.netG, self.netD} do
    local n = ffi.new('int64_t[1]'); check(lib.cenn_trainer_bn_stat_count(self.h, idx - 1, n))
    local st = torch.FloatTensor(tonumber(n[0])); check(lib.cenn_trainer_get_bn_stats_host(self.h, idx - 1, st:data()))
    local off = 1
    for _, m in ipairs(net:findModules('nn.SpatialBatchNormalization')) do
      local C = m.running_mean:nElement()
      m.running_mean:copy(st:narrow(1, off, C)); m.running_var:copy(st:narrow(1, off + C, C)); off = off + 2 * C
    end
  end
end

-- eval-mode generator + full-frame sweep (test_vid_wholeim.lua:98-226, demo.lua:68, test.lua:92)
local Inpainter = torch.class('cenn.Inpainter')
function Inpainter:__init(opt, net, batch)
  local video = opt.predLen ~= nil or opt.inputLen ~= nil
  local cfg = ffi.new('cenn_inpainter_config', {video and 1 or 0, batch or 64, opt.fineSize, opt.nBottleneck, opt.nef or 64, opt.ngf or 64, opt.nc or 3, opt.inputLen or opt.predLen or 1})
  local h = ffi.new('cenn_inpainter*[1]'); check(lib.cenn_inpainter_create(S(), cfg, h)); self.h = ffi.gc(h[0], lib.cenn_inpainter_destroy)
  check(lib.cenn_inpainter_load_host(self.h, net:getParameters():float():data(), bn_stats_of(net):data()))
  self.maskValue, self.F = opt.maskValue, opt.fineSize
end
function Inpainter:forward(x)          -- x: FloatTensor [n, nc*inputLen, F, F] -> same shape (video) or [n, nc, F/2, F/2] (image)
  local out = x.new():resizeAs(x)
  check(lib.cenn_inpainter_forward_host(self.h, x:data(), out:data(), x:size(1))); return out
end
function Inpainter:sweep(images01, mask, init)   -- images01 [P, nc, inh, inw] in [0,1], mask ByteTensor [inh, inw] -> outImages, fullImages, inpaintImages
  local P, nc, inh, inw = images01:size(1), images01:size(2), images01:size(3), images01:size(4)
  local outh, outw = math.ceil(inh / self.F) * self.F, math.ceil(inw / self.F) * self.F
  local o, f, i = torch.FloatTensor(P, nc, outh, outw), torch.FloatTensor(P, nc, outh, outw), torch.FloatTensor(P, nc, outh, outw)
  check(lib.cenn_inpainter_sweep_host(self.h, init and init.h or nil, images01:data(), mask:data(), P, inh, inw, self.maskValue, o:data(), f:data(), i:data()))
  return o, f, i
end

package.loaded['cunn'] = cenn                         -- `require 'cunn'` (train.lua:249) resolves to this module
return cenn
