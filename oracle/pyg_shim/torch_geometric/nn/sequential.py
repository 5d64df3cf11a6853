import torch


def _names(s):
    return [t.strip() for t in s.split(",") if t.strip()]


class Sequential(torch.nn.Module):
    """gnn.Sequential('a, b, c', [(callable, 'a, b -> d'), ...]): named-variable dataflow.
    Child i is registered as `module_{i}` (i = position in the list); plain callables get no
    state_dict key. Returns the output of the last entry."""

    def __init__(self, input_args, modules):
        super().__init__()
        self._inputs = _names(input_args)
        self._steps = []
        for idx, entry in enumerate(modules):
            if isinstance(entry, (tuple, list)):
                fn, desc = entry
                ins, outs = desc.split("->")
                ins, outs = _names(ins), _names(outs)
            else:
                fn, ins, outs = entry, None, None
            name = f"module_{idx}"
            if isinstance(fn, torch.nn.Module):
                self.add_module(name, fn)
            else:
                object.__setattr__(self, name, fn)
            self._steps.append((name, ins, outs))

    def forward(self, *args, **kwargs):
        env = dict(zip(self._inputs, args))
        env.update(kwargs)
        last = args[0] if args else None
        for name, ins, outs in self._steps:
            fn = getattr(self, name)
            if ins is None:
                last = fn(last)
                continue
            last = fn(*[env[k] for k in ins])
            if len(outs) == 1:
                env[outs[0]] = last
            else:
                for k, v in zip(outs, last):
                    env[k] = v
        return last
