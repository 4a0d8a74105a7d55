"""Deferred tensor expressions for outputs that the reference returns as GRAPH NODES (evaluated only when fetched through
sess.run) and that are N*K-sized: e.g. `log_r_nk = tf.log(r_nk)` of gmm.inference / smm.inference (gmm.py:268, smm.py:244).
Evaluating them eagerly would add a full pass over [N,K] to every sweep."""
import torch


class LazyTensor(object):
    """Materialises `fn()` on first use; behaves like the resulting tensor in torch functions, indexing and attribute access."""

    def __init__(self, fn):
        self._fn, self._v = fn, None

    def value(self):
        if self._v is None:
            self._v, self._fn = self._fn(), None
        return self._v

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        un = lambda a: a.value() if isinstance(a, LazyTensor) else a
        args = tuple(un(a) for a in args)
        kwargs = {k: un(v) for k, v in (kwargs or {}).items()}
        return func(*args, **kwargs)

    def __getattr__(self, name):
        return getattr(self.value(), name)

    def __getitem__(self, idx):
        return self.value()[idx]

    def __len__(self):
        return len(self.value())

    def __repr__(self):
        return 'LazyTensor(%r)' % (self._v if self._v is not None else '<deferred>',)

    def __array__(self, dtype=None):
        a = self.value().detach().cpu().numpy()
        return a.astype(dtype) if dtype is not None else a
    for _op in ('add', 'sub', 'mul', 'truediv', 'neg', 'radd', 'rsub', 'rmul', 'rtruediv', 'lt', 'gt', 'le', 'ge'):
        exec("def __%s__(self, *a): return getattr(self.value(), '__%s__')(*a)" % (_op, _op))
    del _op


def lazy_log(t):
    return LazyTensor(lambda: torch.log(t))
