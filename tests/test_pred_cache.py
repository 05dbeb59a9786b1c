"""Prediction cache: on-disk format and axis convention of the reference
(lightning_loftr.py:348-360 writer, test_streetlearn_interiornet.py:250-267 reader), restated with numpy here."""
import os

import numpy as np
import torch

from far_b200 import pred_cache


def _reference_reader(parent, i):
    """The reference's reader, verbatim arithmetic (numpy float64)."""
    pred_path = os.path.join(parent, 'test', 'loftr_preds', str(i) + '.pt')
    num_corr_path = os.path.join(parent, 'test', 'loftr_num_correspondences', str(i) + '.pt')
    if os.path.exists(pred_path) and os.path.exists(num_corr_path):
        loftr_preds = torch.load(pred_path).unsqueeze(0)
        loftr_num_corr = torch.load(num_corr_path).unsqueeze(0)
        loftr_preds = loftr_preds.cpu().numpy()[0]
        T = np.eye(4)
        T[:3] = loftr_preds
        flip_axis = np.array([[1, 0, 0, 0], [0, -1, 0, 0], [0, 0, -1, 0], [0, 0, 0, 1]])
        T = flip_axis @ T @ np.linalg.inv(flip_axis)
        flip_axis = np.array([[0, 1, 0, 0], [1, 0, 0, 0], [0, 0, -1, 0], [0, 0, 0, 1]])
        T = flip_axis @ T @ np.linalg.inv(flip_axis)
        return torch.from_numpy(T.astype(np.double)).unsqueeze(0), loftr_num_corr
    return torch.eye(4)[:3].unsqueeze(0), torch.tensor([0])


def _random_pose(g):
    q, _ = np.linalg.qr(g.standard_normal((3, 3)))
    if np.linalg.det(q) < 0:
        q[:, 0] = -q[:, 0]
    return q, g.standard_normal(3)


def test_roundtrip_matches_reference_reader(tmp_path):
    g = np.random.default_rng(3)
    parent = str(tmp_path)
    poses, counts = [], []
    for i in range(5):
        R, t = _random_pose(g)
        n = int(g.integers(0, 1000))
        pred_cache.save_prediction(parent, 'test', i, R, t, n)
        poses.append(np.concatenate([R, t[:, None]], 1))
        counts.append(n)
    for i in range(5):
        stored = torch.load(os.path.join(parent, 'test', 'loftr_preds', f'{i}.pt'))
        assert stored.dtype == torch.float64 and stored.shape == (3, 4)      # the writer's format
        assert np.array_equal(stored.numpy(), poses[i])
        T_ref, nc_ref = _reference_reader(parent, i)
        T, nc = pred_cache.load_prediction(parent, 'test', i)
        assert T.dtype == torch.float64 and T.shape == (1, 4, 4)
        assert torch.equal(T, T_ref) and torch.equal(nc, nc_ref) and int(nc) == counts[i]
    # missing pair: identity [1,3,4] and zero correspondences, like the reference
    T, nc = pred_cache.load_prediction(parent, 'test', 99)
    T_ref, nc_ref = _reference_reader(parent, 99)
    assert torch.equal(T, T_ref) and torch.equal(nc, nc_ref)


def test_batch_and_in_memory_paths(tmp_path):
    g = np.random.default_rng(4)
    parent = str(tmp_path)
    poses = torch.from_numpy(np.stack([np.concatenate([R, t[:, None]], 1) for R, t in (_random_pose(g) for _ in range(4))]))
    counts = torch.tensor([10, 0, 500, 999])
    pred_cache.save_batch(parent, 'test', [0, 1, 2, 3], poses.float(), counts)   # float32 in, float64 on disk
    T, nc = pred_cache.load_batch(parent, 'test', [0, 1, 2, 3, 7])
    assert T.shape == (5, 4, 4) and T.dtype == torch.float64 and nc.tolist() == [10, 0, 500, 999, 0]
    assert torch.equal(T[4], torch.eye(4, dtype=torch.float64))
    # in-memory path == disk path (up to the float32 the poses were saved from)
    T_mem = pred_cache.to_vit_convention(poses.float())
    assert torch.allclose(T_mem, T[:4], atol=0, rtol=0)
    # conjugation by axis flips keeps a rigid transform rigid and maps t -> (t_y, -t_x ... ) permutation/sign only
    Rm = T_mem[:, :3, :3]
    assert torch.allclose(Rm @ Rm.transpose(1, 2), torch.eye(3, dtype=torch.float64).expand(4, 3, 3), atol=1e-6)
    assert torch.allclose(torch.linalg.det(Rm), torch.ones(4, dtype=torch.float64), atol=1e-6)
    # the in-memory path feeds ViTEss what the disk path would: poses = loftr_rt, count = AFTER-RANSAC inliers
    # (lightning_loftr.py:356-359), not the pre-RANSAC match count
    out = {'loftr_rt': poses.float(), 'num_matches': counts + 1000, 'num_inliers': counts}
    lp, n = pred_cache.from_pipeline(out)
    assert torch.equal(lp, T_mem) and n.dtype == torch.int64 and torch.equal(n, counts)
    pred_cache.save_pipeline(parent, 'mem', [0, 1, 2, 3], out)
    T2, nc2 = pred_cache.load_batch(parent, 'mem', [0, 1, 2, 3])
    assert torch.equal(T2, lp) and torch.equal(nc2, n)


def test_ransac_host_glue_and_sampling_contract(golden_dir):
    """Host-side pieces of the RANSAC round that need no GPU: prior normalisation (setup_prior, ransac.py:176-186), the
    oracle's bias weights against the fixture produced by the unmodified reference (ransac.py:358-367), and the
    counter-based sampling contract (Philox4x32-10 known answers of the Random123 distribution; distinct indices;
    draws follow the weights)."""
    from far_b200.ransac import normalise_prior
    from oracle import far_oracle as O
    g = np.load(os.path.join(golden_dir, "ransac.npz"))
    kp1, kp2 = torch.from_numpy(g["kp1"]), torch.from_numpy(g["kp2"])
    prior = torch.from_numpy(g["prior_rt"])[None].clone()
    prior[:, :, 3] *= 3.7                                   # un-normalised on purpose
    pn = normalise_prior(prior)
    assert torch.allclose(pn[0, :, 3].norm(), torch.tensor(1.0), atol=1e-6)
    assert torch.allclose(pn[0], torch.from_numpy(g["prior_rt"]), atol=1e-6)
    bw = O.ransac_bias_weight(kp1, kp2, pn[0], 0.1)
    assert (bw - torch.from_numpy(g["bias_ref"])).abs().max() < 1e-5
    assert [int(x) for x in O.philox4x32_10(0, 0)] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert [int(x) for x in O.philox4x32_10(0xffffffffffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff)] == \
        [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    w = bw.double().numpy() + 1e-4
    idx = O.ransac_sample_indices(w, 512, seed=3)
    assert idx.min() >= 0 and idx.max() < w.shape[0] and all(len(set(r.tolist())) == 8 for r in idx)
    mass = np.bincount(idx.reshape(-1), minlength=w.shape[0]) / idx.size
    top = np.argsort(-w)[: w.shape[0] // 10]
    assert abs(mass[top].sum() - w[top].sum() / w.sum()) < 0.05
    assert (O.ransac_sample_indices(np.ones(5), 4, seed=3) == -1).all()


def test_mapfree_submission_writer(tmp_path):
    """far_b200.submission against the reference's format (mapfree_6dreg/submission.py:30-82): quaternion w >= 0 that
    reproduces the Gram-Schmidt rotation of the 6-D output, 6-decimal floats, int inliers / '0.0' fallback, NaN poses
    skipped, one file per scene in the zip, no trailing newline."""
    import zipfile
    from far_b200.submission import SubmissionWriter, matrix_to_quaternion_wxyz
    from far_b200.mapfree import rotation_6d_to_matrix
    g = np.random.default_rng(12)
    R6 = torch.from_numpy(g.standard_normal((5, 6)))
    t = torch.from_numpy(g.standard_normal((5, 3)))
    t[3, 1] = float("nan")
    inl = torch.tensor([[120., 30., 2.], [0., 0., 0.], [7., 1., 0.], [5., 0., 0.], [999., 9., 9.]])
    w = SubmissionWriter()
    w.add_batch(["s1", "s1", "s2", "s2", "s1"], [f"seq1/frame_{i:05d}.jpg" for i in range(5)], R6, t, inl)
    path = w.write(str(tmp_path / "sub.zip"))
    with zipfile.ZipFile(path) as z:
        assert sorted(z.namelist()) == ["pose_s1.txt", "pose_s2.txt"]
        s1 = z.read("pose_s1.txt").decode()
        s2 = z.read("pose_s2.txt").decode()
    assert not s1.endswith("\n") and len(s1.split("\n")) == 3 and len(s2.split("\n")) == 1     # frame 3 (NaN) skipped
    R = rotation_6d_to_matrix(R6)
    for line, i in zip(s1.split("\n") + s2.split("\n"), (0, 1, 4, 2)):
        f = line.split(" ")
        assert f[0] == f"seq1/frame_{i:05d}.jpg" and len(f) == 9
        q = np.array([float(x) for x in f[1:5]])
        assert q[0] >= 0 and abs(np.linalg.norm(q) - 1) < 1e-5 and all(len(x.split(".")[1]) == 6 for x in f[1:8])
        qw, qx, qy, qz = q
        Rq = np.array([[1 - 2 * (qy * qy + qz * qz), 2 * (qx * qy - qz * qw), 2 * (qx * qz + qy * qw)],
                       [2 * (qx * qy + qz * qw), 1 - 2 * (qx * qx + qz * qz), 2 * (qy * qz - qx * qw)],
                       [2 * (qx * qz - qy * qw), 2 * (qy * qz + qx * qw), 1 - 2 * (qx * qx + qy * qy)]])
        assert np.abs(Rq - R[i].numpy()).max() < 1e-5
        assert np.abs(np.array([float(x) for x in f[5:8]]) - t[i].numpy()).max() < 1e-6
        assert f[8] == ("0.0" if i == 1 else str(int(inl[i, 0])))
    # the quaternion routine on the four branch cases (trace > 0 and each diagonal dominant)
    Rs = torch.stack([torch.eye(3), torch.diag(torch.tensor([1., -1, -1])), torch.diag(torch.tensor([-1., 1, -1])),
                      torch.diag(torch.tensor([-1., -1, 1]))]).double()
    q = matrix_to_quaternion_wxyz(Rs)
    assert torch.allclose(q.abs(), torch.eye(4, dtype=torch.float64), atol=1e-12)
