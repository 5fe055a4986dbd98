"""TF-checkpoint interchange of the actor weights (agents/agent.py:289-294): table / proto encoding round trip, and the
Keras attribute paths against the ones recorded from the UNMODIFIED reference classes (tests/golden/model_*.npz).
The files could not be opened with TensorFlow itself (not installable here): that part is unpinned."""
import os
import re

import numpy as np
import pytest

from oracle import model as om
from tests import golden_model_util as gu
from tpc_b200 import tf_checkpoint as tfc


def test_crc32c_known_answers():
    assert tfc.crc32c(b"123456789") == 0xE3069283                    # RFC 3720 check value
    assert tfc.crc32c(b"\x00" * 32) == 0x8A9136AA
    assert tfc.masked_crc(b"") == (((0 >> 15) | (0 << 17)) + 0xA282EAD8) & 0xFFFFFFFF


@pytest.mark.parametrize("case", ["rv3_6x10_greedy", "rv3_6x10_cost", "kv2_5x12"])
def test_keras_paths_equal_the_references_attribute_paths(case):
    z, meta = gu.load(case)
    a, _ = gu.net_cfgs(meta)
    ours = [tfc.keras_path(n) for n, _ in om.actor_spec(a)]
    # the generator walked `Sequential.layers[i]`; Keras names those checkpoint edges `layer_with_weights-i`
    ref = [re.sub(r"/layers/(\d+)", r"/layer_with_weights-\1", p.replace(".", "/")) for p in z["actor_paths"]]
    assert ours == ref


def test_round_trip_and_corruption_detection(tmp_path):
    a = om.NetCfg(num_layers=1, dff=128)
    spec = om.actor_spec(a)
    flat = om.golden_params(spec, 5)
    offs, _ = om.spec_offsets(spec)
    tensors = {tfc.keras_path(n): flat[o:o + int(np.prod(s))].reshape(s) for n, (o, s) in offs.items()}
    prefix = str(tmp_path / "actor")
    tfc.write_checkpoint(prefix, tensors)
    assert sorted(os.listdir(tmp_path)) == ["actor.data-00000-of-00001", "actor.index"]
    back = tfc.read_checkpoint(prefix)
    assert set(back) == set(tensors) and all(np.array_equal(back[k], tensors[k]) for k in tensors)
    assert os.path.getsize(prefix + ".data-00000-of-00001") >= flat.nbytes
    raw = bytearray(open(prefix + ".data-00000-of-00001", "rb").read())
    raw[-100] ^= 0xFF                                   # inside the last float tensor
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(raw))
    with pytest.raises(ValueError):
        tfc.read_checkpoint(prefix)
    idx = bytearray(open(prefix + ".index", "rb").read())
    idx[20] ^= 0xFF
    open(prefix + ".index", "wb").write(bytes(idx))
    with pytest.raises(ValueError):
        tfc.read_checkpoint(prefix)


def test_object_graph_lists_every_variable():
    paths = ["encoder/embedding1/layer/kernel", "encoder/embedding1/layer/bias", "decoder/last_decoder_layer/V/kernel"]
    g = tfc._object_graph(paths)
    nodes = [v for n, _, v in tfc._parse(g) if n == 1]
    keys = []
    for nd in nodes:
        for n, _, v in tfc._parse(nd):
            if n == 2:
                keys += [vv.decode() for nn, _, vv in tfc._parse(v) if nn == 3]
    assert sorted(keys) == sorted(p + tfc.VAR_SUFFIX for p in paths)
    root_children = [dict((nn, vv) for nn, _, vv in tfc._parse(v))[2].decode() for n, _, v in tfc._parse(nodes[0]) if n == 1]
    assert root_children == ["encoder", "decoder"]
