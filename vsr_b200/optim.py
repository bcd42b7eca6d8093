"""FlatAdam — torch.optim.Adam semantics as ONE fused kernel over the net's flat fp32 buckets
(reference: the optimizer is `torch.optim.<name>(net.parameters(), **kwargs)`, main.py:73-74, stepped
in acdc_vsr_trainer.py:46).  `grad_scale` folds the 1/world_size of the data-parallel mean."""
import torch


class FlatAdam(torch.optim.Optimizer):
    _RING = 4      # pinned hyper-parameter buffers in flight (the host may run this many steps ahead of the device)

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False, net=None):
        params = list(params)
        if amsgrad:
            raise NotImplementedError("FlatAdam: amsgrad is not implemented by the fused kernel (use torch.optim.Adam)")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        if len(self.param_groups) != 1:
            raise NotImplementedError("FlatAdam optimises ONE parameter group (the net's flat bucket); got "
                                      f"{len(self.param_groups)} groups")
        self._net = net
        self._m = self._v = None
        self._step = 0
        self.grad_scale = 1.0
        self._hyper_host, self._hyper_ev, self._hyper_dev, self._hyper_i = [], [], None, 0

    def bind(self, net):
        """bind to a vsr_b200 net (its parameters must be exactly this optimizer's parameters)."""
        self._net = net
        return self

    def _ensure_state(self):
        net = self._net
        if net is None:
            raise RuntimeError("FlatAdam.bind(net) must be called with the vsr_b200 net it optimises")
        if not net._is_flat():
            net._flatten()
        flat = net.flat
        if self._m is None or self._m.device != flat.device or self._m.numel() != flat.numel():
            m_old, v_old = self._m, self._v
            self._m, self._v = torch.zeros_like(flat), torch.zeros_like(flat)
            if m_old is not None and m_old.numel() == flat.numel():
                self._m.copy_(m_old)
                self._v.copy_(v_old)
            for p, ref in zip(net.parameters(), net._plan.params.values()):
                n = p.numel()
                self.state[p] = {"step": torch.tensor(float(self._step)),
                                 "exp_avg": self._m[ref.offset:ref.offset + n].view(ref.shape),
                                 "exp_avg_sq": self._v[ref.offset:ref.offset + n].view(ref.shape)}
        return flat

    def prepare_step(self):
        """host side of a step: advance the step count and upload {lr, betas, eps, wd, step, grad_scale}
        to the device (outside any captured graph)."""
        flat = self._ensure_state()
        grp = self.param_groups[0]
        self._step += 1
        vals = [float(grp["lr"]), grp["betas"][0], grp["betas"][1], grp["eps"], grp["weight_decay"],
                float(self._step), self.grad_scale]
        if len(self.param_groups) != 1:
            raise NotImplementedError("FlatAdam optimises ONE parameter group (add_param_group is not supported)")
        if self._hyper_dev is None or self._hyper_dev.device != flat.device:
            self._hyper_dev = torch.zeros(7, dtype=torch.float32, device=flat.device)
            self._hyper_host = [torch.zeros(7, dtype=torch.float32) for _ in range(self._RING)]
            self._hyper_ev = [None] * self._RING
            if flat.is_cuda:
                self._hyper_host = [h.pin_memory() for h in self._hyper_host]
            self._hyper_np = [h.numpy() for h in self._hyper_host]     # views: filled without building a tensor per step
        # The H2D copy below is asynchronous and the host runs ahead of the device (graph replay, one sync per epoch):
        # a pinned buffer is rewritten only after the copy that last read it has executed (ring + events), otherwise
        # step k could see the step count / lr of step k+1.
        i = self._hyper_i
        self._hyper_i = (i + 1) % self._RING
        if self._hyper_ev[i] is not None:
            self._hyper_ev[i].synchronize()
        self._hyper_np[i][:] = vals
        self._hyper_dev.copy_(self._hyper_host[i], non_blocking=True)
        if flat.is_cuda:
            ev = self._hyper_ev[i] or torch.cuda.Event()
            ev.record(torch.cuda.current_stream(flat.device))
            self._hyper_ev[i] = ev

    def launch(self, flat_grad):
        """device side of a step: one fused kernel (graph-capturable)."""
        flat = self._net.flat
        self._net._backend().adam_flat_dev(flat, flat_grad.to(flat.dtype) if flat_grad.dtype != flat.dtype else flat_grad,
                                           self._m, self._v, self._hyper_dev.to(flat.dtype) if flat.dtype != torch.float32 else self._hyper_dev)

    @torch.no_grad()
    def step(self, closure=None, flat_grad=None):
        loss = closure() if closure is not None else None
        self._ensure_state()
        g = flat_grad if flat_grad is not None else self._net.flat_grad
        if g is None:
            raise RuntimeError("FlatAdam.step: no flat gradient (run backward first)")
        self.prepare_step()
        self.launch(g)
        return loss

    def state_dict(self):
        if self._net is not None and self._m is not None:
            for st in self.state.values():
                st["step"] = torch.tensor(float(self._step))
        return super().state_dict()

    def load_state_dict(self, sd):
        super().load_state_dict(sd)
        steps = [float(st["step"]) for st in self.state.values() if "step" in st]
        self._step = int(steps[0]) if steps else 0
        if self._net is not None:
            loaded = {p: dict(st) for p, st in self.state.items()}
            self._m = None
            flat = self._ensure_state()
            for p in self._net.parameters():
                if p in loaded and "exp_avg" in loaded[p]:
                    self.state[p]["exp_avg"].copy_(loaded[p]["exp_avg"].to(flat.device))
                    self.state[p]["exp_avg_sq"].copy_(loaded[p]["exp_avg_sq"].to(flat.device))
