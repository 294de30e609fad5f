"""stdout logging with the reference's line format (utils/log_utils.py:34-56); no matplotlib."""


def log_values(cost, grad_norms, epoch, batch_id, step, log_likelihood, reinforce_loss, bl_loss, tb_logger, opts):
    avg_cost = cost.mean().item()
    grad_norms, grad_norms_clipped = grad_norms
    print('epoch: {}, train_batch_id: {}, avg_cost: {}'.format(epoch, batch_id, avg_cost))
    print('loss: {}, grad norm: {}, clipped: {}'.format(reinforce_loss, grad_norms[0], grad_norms_clipped[0]))
    if not opts.no_tensorboard and tb_logger is not None:
        tb_logger.log_value('avg_cost', avg_cost, step)
        tb_logger.log_value('actor_loss', reinforce_loss.item(), step)
        tb_logger.log_value('nll', -log_likelihood.mean().item(), step)
        tb_logger.log_value('grad_norm', grad_norms[0], step)
        tb_logger.log_value('grad_norm_clipped', grad_norms_clipped[0], step)
