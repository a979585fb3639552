import sys; sys.path.insert(0, '/root/repo')
import numpy as np, torch
from graphgan_b200.discriminator import Discriminator
from oracle import updates
dev = torch.device('cuda:0')
for (n, d, B) in [(400, 256, 128)]:
    rs = np.random.RandomState(n + d)
    print('CONFIG', n, d, B)
    emb = rs.normal(0, 0.5, size=(n, d))
    dm = Discriminator(n, emb, device=dev); ora = updates.Discriminator(n, emb, 1e-3, 1e-5)
    for step in range(6):
        i, j = rs.randint(0, n, B).astype(np.int32), rs.randint(0, n, B).astype(np.int32)
        lab = (rs.random_sample(B) < 0.5).astype(np.float32)
        rows, g_rows, g_bias = ora.grads(i, j, lab)
        # device grads
        import ctypes as C
        from graphgan_b200 import _cabi
        ti, tj, ta = dm._dev_i32(i), dm._dev_i32(j), dm._dev_f32(lab)
        _cabi.check(dm.lib.gg_pair_grad(0, B, 0, ti.data_ptr(), tj.data_ptr(), ta.data_ptr(), dm.emb.data_ptr(), dm.bias_t.data_ptr(), dm.ld, C.c_float(1e-5), dm.n_unique.data_ptr(), dm.uniq_ids.data_ptr(), dm.grad_rows.data_ptr(), dm.grad_bias.data_ptr(), dm.row_slot.data_ptr(), 0), 'g')
        torch.cuda.synchronize()
        U = int(dm.n_unique.item()); ids = dm.uniq_ids[:U].cpu().numpy(); gr = dm.grad_rows[:U, :d].cpu().numpy(); gb = dm.grad_bias[:U].cpu().numpy()
        print(n, d, B, 'step', step, 'U', U, len(rows), 'ids equal', np.array_equal(ids, rows), 'grad maxabs diff', np.abs(gr - g_rows).max(), 'rel', (np.abs(gr - g_rows) / (np.abs(g_rows) + 1e-12)).max(), 'bias diff', np.abs(gb - g_bias).max())
        dm.apply_adam(); ora.adam.apply(ora.E, ora.b, rows, g_rows, g_bias)
        E = dm.embedding_numpy(); dif = np.abs(E - ora.E); rel = dif / (np.abs(ora.E) + 1e-30)
        k = np.unravel_index(np.argmax(dif), dif.shape)
        print('   fro', np.linalg.norm(E.astype(np.float64)-ora.E)/np.linalg.norm(ora.E), 'okfrac', float(np.mean(dif <= 1e-5*np.abs(ora.E)+1e-6)), 'n>1e-5', int((dif>1e-5).sum()), 'n>1e-4', int((dif>1e-4).sum()))
        print('   emb maxabs', dif.max(), 'at', k, E[k], ora.E[k], 'viol', int((dif > 1e-5 * np.abs(ora.E) + 2e-7).sum()), 'm diff', np.abs(dm.m_emb[:, :d].cpu().numpy() - ora.adam.m_e).max(), 'v rel', (np.abs(dm.v_emb[:, :d].cpu().numpy() - ora.adam.v_e) / (ora.adam.v_e + 1e-30)).max())
