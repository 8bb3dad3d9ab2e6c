"""Model base class -- same contract as the reference's
``mbrl_traffic/models/base.py:4-129``."""


class Model(object):
    """Base model object.

    The model estimates next-step observations from current observations and
    actions.  Attributes and method names mirror the reference
    (``models/base.py:26-129``) so the classes in this package drop in under
    ``algorithm.py:279-286`` and ``evaluate_model.py:369-403``.

    Attributes
    ----------
    sess, ob_space, ac_space, replay_buffer, verbose
        stored as given; ``sess`` (a TensorFlow session in the reference) is
        never used by this package.
    """

    def __init__(self, sess, ob_space, ac_space, replay_buffer, verbose):
        self.sess = sess
        self.ob_space = ob_space
        self.ac_space = ac_space
        self.replay_buffer = replay_buffer
        self.verbose = verbose

    def initialize(self):
        """Initialize the model (base.py:50-56)."""
        raise NotImplementedError

    def get_next_obs(self, obs, action):
        """Estimate the next-step observation (base.py:58-74)."""
        raise NotImplementedError

    def update(self):
        """Perform a gradient/parameter update step (base.py:76-84)."""
        raise NotImplementedError

    def compute_loss(self, states, actions, next_states):
        """Compute the loss for a batch of data (base.py:86-102)."""
        raise NotImplementedError

    def get_td_map(self):
        """Return dict map for the summary (base.py:104-109)."""
        raise NotImplementedError

    def save(self, save_path):
        """Save the parameters of the model (base.py:111-119)."""
        raise NotImplementedError

    def load(self, load_path):
        """Load model parameters (base.py:121-129)."""
        raise NotImplementedError
