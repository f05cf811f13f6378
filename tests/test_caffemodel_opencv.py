"""A THIRD-PARTY reader of the Caffe formats and implementation of the Caffe layer semantics this path uses: OpenCV's
Caffe importer (`cv2.dnn.readNetFromCaffe`, OpenCV's own protobuf + caffe.proto + layer code, no line of it written here).

The reference's arithmetic lives in BVLC Caffe, which cannot be installed in this image, so nothing upstream pins the
oracle.  This test pins what a third party CAN check without a GPU:
  * the `.caffemodel` files of the host mirror (`host/caffe_proto.cpp`, wire format written by hand) are NetParameter
    files a real Caffe parser reads: layer names, blob shapes and data arrive intact;
  * the forward semantics the oracle restates - InnerProduct `y = x W^T + b` with W stored [out x in], in-place leaky ReLU
    with negative_slope 0.01, Concat of (states, actions, action_params) along axis 2, the two linear actor heads in the
    order action_layer(4), actionpara_layer(6) - against OpenCV's implementation of those layers, on the reference's own
    layer names and blob layout ([N,1,S,1] MemoryData blobs, dqn.cpp:400-454).
"""
import os
import subprocess

import numpy as np
import pytest

from util import ROOT
from oracle import oracle as O

cv2 = pytest.importorskip("cv2")
HOST = os.path.join(ROOT, "dqn-hfo_b200", "host")


def deploy_prototxt(S, hidden, critic, n):
    """Deploy form of CreateActorNet / CreateCriticNet (dqn.cpp:418-454): the MemoryData layers become net inputs."""
    t = [f'name: "{"Critic" if critic else "Actor"}"', f'input: "states" input_shape {{ dim: {n} dim: 1 dim: {S} dim: 1 }}']
    bottom = "states"
    if critic:
        t += [f'input: "actions" input_shape {{ dim: {n} dim: 1 dim: 4 dim: 1 }}',
              f'input: "action_params" input_shape {{ dim: {n} dim: 1 dim: 6 dim: 1 }}',
              'layer { name: "concat" type: "Concat" bottom: "states" bottom: "actions" bottom: "action_params" top: "state_actions" '
              'concat_param { axis: 2 } }']
        bottom = "state_actions"
    for i, h in enumerate(hidden, 1):
        t.append(f'layer {{ name: "ip{i}_layer" type: "InnerProduct" bottom: "{bottom}" top: "ip{i}" inner_product_param {{ num_output: {h} }} }}')
        t.append(f'layer {{ name: "ip{i}_relu_layer" type: "ReLU" bottom: "ip{i}" top: "ip{i}" relu_param {{ negative_slope: 0.01 }} }}')
        bottom = f"ip{i}"
    if critic:
        t.append(f'layer {{ name: "q_values_layer" type: "InnerProduct" bottom: "{bottom}" top: "q_values" inner_product_param {{ num_output: 1 }} }}')
    else:
        t.append(f'layer {{ name: "action_layer" type: "InnerProduct" bottom: "{bottom}" top: "actions" inner_product_param {{ num_output: 4 }} }}')
        t.append(f'layer {{ name: "actionpara_layer" type: "InnerProduct" bottom: "{bottom}" top: "action_params" inner_product_param {{ num_output: 6 }} }}')
    return "\n".join(t) + "\n"


@pytest.mark.parametrize("S,hidden", [(59, (1024, 512, 256, 128)), (77, (96, 64, 48, 32)), (58, (64, 32))])
def test_opencv_reads_our_caffemodel_and_agrees_with_the_oracle_forward(tmp_path, S, hidden):
    subprocess.run(["make", "-C", HOST, "host_selftest"], check=True, stdout=subprocess.DEVNULL)
    tool = os.path.join(HOST, "host_selftest")
    n = 7
    rng = np.random.default_rng(S + len(hidden))
    cfg = O.make_config(state_size=S, batch=n, hidden=hidden)
    actor, critic = O.init_params(cfg, False, rng, "warm"), O.init_params(cfg, True, rng, "warm")
    st = O.OracleState(cfg, actor, critic, actor, critic)
    s, a10 = rng.uniform(-1, 1, (n, S)).astype(np.float32), rng.uniform(-1, 1, (n, 10)).astype(np.float32)
    a10[:, 4:] *= 50.0
    hid = ",".join(str(h) for h in hidden)
    for is_critic, flat in ((False, actor), (True, critic)):
        kind = "critic" if is_critic else "actor"
        wbin, model, proto = tmp_path / f"{kind}.bin", tmp_path / f"{kind}.caffemodel", tmp_path / f"{kind}.prototxt"
        flat.astype(np.float32).tofile(wbin)
        subprocess.run([tool, "--write-caffemodel", str(model), kind, str(S), hid, str(wbin)], check=True)
        proto.write_text(deploy_prototxt(S, hidden, is_critic, n))
        net = cv2.dnn.readNetFromCaffe(str(proto), str(model))            # third-party parser of both files
        # every parametrised layer arrived with the blobs of the flat array, in Caffe's [out x in] / [out] shapes
        off, dims = 0, [S + (10 if is_critic else 0)] + list(hidden)
        names = [f"ip{i}_layer" for i in range(1, len(hidden) + 1)] + (["q_values_layer"] if is_critic else ["action_layer", "actionpara_layer"])
        outs = list(hidden) + ([1] if is_critic else [4, 6])
        ins = dims[:-1] + [hidden[-1]] * (1 if is_critic else 2)
        for name, o, i in zip(names, outs, ins):
            lid = net.getLayerId(name)
            assert lid > 0, name
            W, b = net.getParam(lid, 0), net.getParam(lid, 1)
            assert W.size == o * i and b.size == o, (name, W.shape, b.shape)
            assert np.array_equal(W.reshape(o, i), flat[off:off + o * i].reshape(o, i)), name
            assert np.array_equal(b.reshape(o), flat[off + o * i:off + o * i + o]), name
            off += o * i + o
        assert off == flat.size
        # forward through OpenCV's layers == the oracle's restatement of Caffe's
        net.setInput(s.reshape(n, 1, S, 1), "states")
        if is_critic:
            net.setInput(np.ascontiguousarray(a10[:, :4]).reshape(n, 1, 4, 1), "actions")
            net.setInput(np.ascontiguousarray(a10[:, 4:]).reshape(n, 1, 6, 1), "action_params")
            got = net.forward("q_values_layer").reshape(n)
            ref = st.critic_forward(s, a10).reshape(n)
        else:
            o4, o6 = net.forward(["action_layer", "actionpara_layer"])
            got = np.concatenate([o4.reshape(n, 4), o6.reshape(n, 6)], axis=1)
            ref = st.actor_forward(s)
        err = np.abs(got - ref).max() / np.abs(ref).max()
        assert err < 2e-6, (kind, err)
        # and the reader of the mirror takes the file back bit for bit
        back = tmp_path / f"{kind}_back.bin"
        subprocess.run([tool, "--read-caffemodel", str(model), kind, str(S), hid, str(back)], check=True, stdout=subprocess.DEVNULL)
        assert np.array_equal(np.fromfile(back, np.float32), flat)


def test_opencv_leaky_relu_and_negative_inputs():
    """The slope applies to negative pre-activations only and zero stays zero (Caffe ReLU forward:
    max(x,0) + negative_slope*min(x,0)) - checked on a hand-made 1-layer net whose pre-activations have both signs."""
    import tempfile
    subprocess.run(["make", "-C", HOST, "host_selftest"], check=True, stdout=subprocess.DEVNULL)
    with tempfile.TemporaryDirectory() as d:
        S, hidden, n = 58, (64, 32), 5
        cfg = O.make_config(state_size=S, batch=n, hidden=hidden)
        rng = np.random.default_rng(1)
        actor = O.init_params(cfg, False, rng, "warm")
        actor[:S * 64] = 0.0                                   # ip1 weights zero -> pre-activation = bias only
        bias = np.linspace(-2, 2, 64).astype(np.float32); bias[10] = 0.0
        actor[S * 64:S * 64 + 64] = bias
        actor.tofile(os.path.join(d, "a.bin"))
        subprocess.run([os.path.join(HOST, "host_selftest"), "--write-caffemodel", os.path.join(d, "a.caffemodel"), "actor", str(S), "64,32",
                        os.path.join(d, "a.bin")], check=True)
        open(os.path.join(d, "a.prototxt"), "w").write(deploy_prototxt(S, hidden, False, n))
        net = cv2.dnn.readNetFromCaffe(os.path.join(d, "a.prototxt"), os.path.join(d, "a.caffemodel"))
        net.setInput(rng.uniform(-1, 1, (n, 1, S, 1)).astype(np.float32), "states")
        h1 = net.forward("ip1_relu_layer").reshape(n, 64)
        expect = np.where(bias > 0, bias, np.float32(0.01) * bias)
        assert np.allclose(h1, np.broadcast_to(expect, (n, 64)), rtol=1e-6, atol=0)
        assert (h1[:, 10] == 0).all()


@pytest.mark.parametrize("kind", ["actor", "critic"])
def test_opencv_parses_the_prototxt_the_mirror_writes(tmp_path, kind):
    """`<prefix>_{actor,critic}.prototxt` (dqn_main.cpp:232-246) as written by host/prototxt.cpp goes through OpenCV's
    protobuf text-format parser with the real caffe.proto (MemoryData / InnerProduct / ReLU / Concat / Silence /
    EuclideanLoss layers, `force_backward`, fillers): unknown fields or a malformed message would be a parse error."""
    subprocess.run(["make", "-C", HOST, "host_selftest"], check=True, stdout=subprocess.DEVNULL)
    f = tmp_path / f"{kind}.prototxt"
    subprocess.run([os.path.join(HOST, "host_selftest"), "--write-prototxt", str(f), kind, "59", "1024,512,256,128", "32"], check=True)
    text = f.read_text()
    assert "force_backward: true" in text and 'type: "MemoryData"' in text
    net = cv2.dnn.readNetFromCaffe(str(f))
    names = list(net.getLayerNames())
    want = ["state_input_layer"] + [n for i in range(1, 5) for n in (f"ip{i}_layer", f"ip{i}_relu_layer")]
    want += ["action_input_layer", "action_params_input_layer", "target_input_layer", "concat", "q_values_layer", "loss"] if kind == "critic" \
        else ["action_layer", "actionpara_layer"]
    for n in want:
        assert n in names, (n, names)
    # tower order as in dqn.cpp:400-416
    idx = [names.index(f"ip{i}_layer") for i in range(1, 5)]
    assert idx == sorted(idx)
