"""Device plan maintenance (stargcn_b200.graph.MultiLinkCSR): construction from per-level lists without host
concatenation, and the in-place refresh + rebuild used by same-shaped iterations (CUDA-graph capturable)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def layer_lists(shape="ml-100k", seed=1000):
    from stargcn_b200 import synth
    return synth.make_layer_inputs(shape, seed=seed)


def build(wl, side, **kw):
    from stargcn_b200.graph import MultiLinkCSR
    n_nb = wl["n_item"] if side == "user" else wl["n_user"]
    return MultiLinkCSR(*wl[side][:3], n_nb=n_nb, device="cuda", **kw)


def test_construction_from_numpy_pinned_and_device_lists_agree():
    wl = layer_lists()
    ep_l, ptr_l, sup_l = wl["user"][:3]
    ref = build(wl, "user", validate=True)
    want = np.concatenate(ep_l), np.concatenate(sup_l)
    assert np.array_equal(ref.end_points.cpu().numpy(), want[0]) and np.array_equal(ref.support.cpu().numpy(), want[1])
    offs = np.concatenate([[0], np.cumsum([int(p[-1]) for p in ptr_l])])
    cat = np.concatenate([[0]] + [ptr_l[r][1:].astype(np.int64) + offs[r] for r in range(wl["R"])])
    assert np.array_equal(ref.cat_indptr.cpu().numpy(), cat)
    from stargcn_b200.graph import MultiLinkCSR
    pinned = [[torch.from_numpy(a).pin_memory() for a in lst] for lst in (ep_l, ptr_l, sup_l)]
    on_dev = [[torch.from_numpy(a).cuda() for a in lst] for lst in (ep_l, ptr_l, sup_l)]
    for lists, kw in ((pinned, {}), (on_dev, {}), (on_dev, dict(nnz_l=[int(p[-1]) for p in ptr_l]))):
        c = MultiLinkCSR(*lists, n_nb=wl["n_item"], device="cuda", **kw)
        assert torch.equal(c.end_points, ref.end_points) and torch.equal(c.support, ref.support)
        assert torch.equal(c.cat_indptr, ref.cat_indptr) and c.nnz_l == ref.nnz_l
    # the reference's length-1 dummies for an empty level (graph.py:221-222) and the validation switch
    ep2, sup2, ptr2 = list(ep_l), list(sup_l), list(ptr_l)
    ep2[2], sup2[2], ptr2[2] = np.zeros(1, np.int32), np.zeros(1, np.float32), np.zeros_like(ptr_l[2])
    c = MultiLinkCSR(ep2, ptr2, sup2, n_nb=wl["n_item"], device="cuda", validate=True)
    assert c.nnz == ref.nnz - int(ptr_l[2][-1]) and c.nnz_l[2] == 0
    bad = [e.copy() for e in ep_l]
    bad[1][3] = wl["n_item"]
    with pytest.raises(ValueError):
        MultiLinkCSR(bad, ptr_l, sup_l, n_nb=wl["n_item"], device="cuda", validate=True)


def test_refresh_in_place_and_graph_replay():
    """load_lists_ + rebuild_ on a same-shaped plan: derived structures equal a freshly built plan, and a CUDA graph
    that contains rebuild_ + forward + backward follows new list contents."""
    from stargcn_b200 import runtime
    from stargcn_b200.layers import MultiLinkGCNAggregator
    wl = layer_lists()
    R, D, U = wl["R"], wl["D"], 250
    ep_l, ptr_l, sup_l = wl["user"][:3]
    rs = np.random.RandomState(0)
    # a second plan of the SAME shape: same pattern, permuted end points inside every segment is not needed —
    # new supports and a rotation of the item ids keep every per-level count
    ep_b = [((e.astype(np.int64) + 7) % wl["n_item"]).astype(np.int32) for e in ep_l]
    sup_b = [rs.uniform(0.1, 1.0, s.shape).astype(np.float32) for s in sup_l]
    csr = build(wl, "user").prepare(backward=True)
    fresh_b = type(csr)(ep_b, ptr_l, sup_b, n_nb=wl["n_item"], device="cuda").prepare(backward=True)
    agg = MultiLinkGCNAggregator(units=U, num_links=R, act="leaky", ordinal_sharing=False, accum="sum", in_units=D).cuda()
    x = torch.randn(wl["n_item"], D, device="cuda").requires_grad_(True)
    gout = torch.randn(wl["n_user"], U, device="cuda")
    holder = {}

    def step():
        x.grad = None
        for p in agg.parameters():
            p.grad = None
        csr.rebuild_()
        out = agg(x, csr)
        out.backward(gout)
        holder["out"] = out.detach()

    def reference(plan):
        x2 = x.detach().clone().requires_grad_(True)
        for p in agg.parameters():
            p.grad = None
        o = agg(x2, plan)
        o.backward(gout)
        return o.detach().clone(), x2.grad.clone(), agg.weight1.grad.clone()

    want_a, want_b = reference(build(wl, "user")), reference(fresh_b)
    graphed = runtime.GraphedStep(step)
    graphed(); torch.cuda.synchronize()
    assert torch.equal(holder["out"], want_a[0]) and torch.equal(x.grad, want_a[1]) and torch.equal(agg.weight1.grad, want_a[2])
    pin = lambda lst: [torch.from_numpy(a).pin_memory() for a in lst]
    csr.load_lists_(pin(ep_b), pin(ptr_l), pin(sup_b))
    graphed(); torch.cuda.synchronize()
    assert torch.equal(csr.transposed()[1], fresh_b.transposed()[1]) and torch.equal(csr.transposed()[2], fresh_b.transposed()[2])
    assert torch.equal(holder["out"], want_b[0]) and torch.equal(x.grad, want_b[1]) and torch.equal(agg.weight1.grad, want_b[2])
    with pytest.raises(ValueError):
        csr.load_lists_(pin(ep_b)[:-1], pin(ptr_l)[:-1], pin(sup_b)[:-1])
    # the two transports (one kernel reading the pinned arrays / one DMA copy per list) deliver the same arrays
    a, b = build(wl, "user"), build(wl, "user")
    a.load_lists_(pin(ep_b), pin(ptr_l), pin(sup_b), zero_copy=True)
    b.load_lists_(pin(ep_b), pin(ptr_l), pin(sup_b), zero_copy=False)
    torch.cuda.synchronize()
    assert torch.equal(a.end_points, b.end_points) and torch.equal(a.support, b.support) and torch.equal(a.cat_indptr, b.cat_indptr)
    assert torch.equal(a.end_points, fresh_b.end_points) and torch.equal(a.cat_indptr, fresh_b.cat_indptr)
    with pytest.raises(ValueError):
        a.load_lists_(ep_b, ptr_l, sup_b, zero_copy=True)          # pageable numpy arrays cannot be read by the kernel


def test_upload_segments_ragged_lengths_and_misaligned_slices():
    """sg_upload_segments: 16-byte units with partial tails, device slices at arbitrary 4-byte offsets, host arrays that
    are NOT 16-byte aligned (word-by-word path), empty segments, and more segments than one launch holds."""
    import ctypes
    from stargcn_b200 import _lib
    lib = _lib.load()
    rs = np.random.RandomState(3)
    lens = [1, 2, 3, 4, 5, 7, 8, 0, 31, 32, 33, 1000, 4097] + [int(v) for v in rs.randint(0, 300, size=70)]   # 83 segments
    host = torch.from_numpy(rs.randint(-2 ** 31, 2 ** 31 - 1, size=sum(lens) + 4 * len(lens) + 8, dtype=np.int64).astype(np.int32)).pin_memory()
    dev = torch.full((sum(lens) + 3 * len(lens) + 8,), -7, dtype=torch.int32, device="cuda")
    srcs, dsts, nbytes, want = [], [], [], dev.clone()
    ho, do = 0, 1
    for i, n in enumerate(lens):
        ho += i % 4                         # host sub-array starts at every 4-byte phase of a 16-byte line
        do += (i * 3) % 4 + (1 if i % 5 == 0 else 0)
        srcs.append(host.data_ptr() + 4 * ho); dsts.append(dev.data_ptr() + 4 * do); nbytes.append(4 * n)
        want[do:do + n] = host[ho:ho + n].cuda()
        ho += n; do += n
    n = len(lens)
    _lib.check(lib.sg_upload_segments((ctypes.c_void_p * n)(*dsts), (ctypes.c_void_p * n)(*srcs), (ctypes.c_size_t * n)(*nbytes), n, None),
               "sg_upload_segments")
    torch.cuda.synchronize()
    assert torch.equal(dev, want)           # every word where it belongs, nothing outside the slices touched
