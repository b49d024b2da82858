"""Mirror of the reference's ``utils`` package for the calibration hot path.

``visualize_per_layer`` exists because ``dfq.py:5`` imports it (utils/__init__.py:1-12 in the reference):
a per-output-channel box plot of a weight tensor, only drawn when ``visualize_state=True``.
"""


def visualize_per_layer(param, title='test'):
    import matplotlib.pyplot as plt
    channels = param.detach().cpu().reshape(param.shape[0], -1)
    fig, axis = plt.subplots()
    axis.set_title(title)
    axis.boxplot([row.numpy() for row in channels], showfliers=False)
    plt.show()
