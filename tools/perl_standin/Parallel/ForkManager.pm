package Parallel::ForkManager;
# Minimal stand-in for the CPAN module Parallel::ForkManager, which the
# reference scripts `use` but which is not installed in this image.  Only the
# four methods the reference calls are provided (new/start/finish/
# wait_all_children).  It carries no arithmetic; it exists solely so that
# tools/make_goldens.py can run the UNMODIFIED reference to generate goldens.
use strict;
use warnings;

sub new {
    my ($class, $max) = @_;
    $max = 0 unless defined $max;
    return bless { max => $max, kids => {} }, $class;
}

# max == 0 mimics the real module's "debug" mode: no fork, run inline.
sub start {
    my ($self) = @_;
    return 0 if $self->{max} <= 0;
    while (scalar(keys %{ $self->{kids} }) >= $self->{max}) {
        my $pid = waitpid(-1, 0);
        last if $pid <= 0;
        delete $self->{kids}{$pid};
    }
    my $pid = fork();
    die "fork failed: $!" unless defined $pid;
    if ($pid) { $self->{kids}{$pid} = 1; return $pid; }
    $self->{in_child} = 1;
    return 0;
}

sub finish {
    my ($self) = @_;
    exit 0 if $self->{in_child};
    return 0;
}

sub wait_all_children {
    my ($self) = @_;
    while (scalar(keys %{ $self->{kids} })) {
        my $pid = waitpid(-1, 0);
        last if $pid <= 0;
        delete $self->{kids}{$pid};
    }
}

1;
